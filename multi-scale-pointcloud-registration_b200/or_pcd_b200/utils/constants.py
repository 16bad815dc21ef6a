"""Default parameters of the reference (or_pcd/utils/constants.py:26-81), by value."""
import math

# statistical outlier removal (constants.py:26,30)
NB_NEIGHBOURS = 64
STD_RATIO = 2

# down-samplers (constants.py:38-51)
RANDOM_SAMPLE_SIZE = 15000
FARTHEST_SAMPLE_SIZE = 10000
VOXEL_SAMPLE_SIZE = 2000
BASE_VOXEL_SIZE = 0.01
MIN_VOXEL_SIZE = 0.0001
VOXEL_DELTA = 0.01
VOXEL_EPS = 0.0005
SAMPLE_SIZE = 4096

# fast global registration (constants.py:54-61)
DIVISION_FACTOR = 1.4
TUPLE_SCALE = 0.9
ITERATION_NUMBER = 100
DECREASE_MU = True
NORMAL_ESTIMATE_RADIUS = 0.1
NORMAL_ESTIMATE_KNN = 20
FPFH_RADIUS = 0.1
FPFH_KNN = 20

# generalized ICP (constants.py:64,67)
MAX_ITERATIONS = 100
MAXIMUM_CORRESPONDENCE_DISTANCE = 0.5

# aligner (constants.py:71-77)
MULTISTART_ATTEMPTS = 30
ALIGNER_DEG = math.pi / 2
ALIGNER_MU = 0.0
ALIGNER_STD = 0.1
ALIGNER_MAX_ITER = 100
ALIGNER_DELTA = 0.2
ALIGNER_EPS = 0.05

# refiner (constants.py:80-81)
REFINER_MAX_ITER = 200
REFINER_DISTANCE_THRESHOLD = 0.5

# Open3D defaults behind the reference's calls
GICP_EPSILON = 1e-3
GICP_COVARIANCE_KNN = 20
ICP_RELATIVE_FITNESS = 1e-6
ICP_RELATIVE_RMSE = 1e-6
