from .BaseScaler import BaseScaler
from .radiusScaler import RadiusScaler
