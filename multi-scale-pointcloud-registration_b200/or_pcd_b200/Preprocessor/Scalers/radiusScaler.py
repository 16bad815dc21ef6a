"""Mean-centre and divide by the largest distance to the mean
(reference: Preprocessor/Scalers/radiusScaler.py:18-30).

O(N) host glue that fixes the float64 inputs of everything downstream, so it is
kept in numpy with the reference's exact operations (np.mean, and the euclidean
norm summed in coordinate order like scipy's cdist) to stay bit-identical."""
import numpy as np

from .BaseScaler import BaseScaler


class RadiusScaler(BaseScaler):
    def process(self, cloud: np.ndarray) -> np.ndarray:
        center = np.mean(cloud, axis=0, keepdims=True)
        d = cloud - center
        radius = np.max(np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]))
        self.mean = center
        self.scale = radius
        return (cloud - center) / radius
