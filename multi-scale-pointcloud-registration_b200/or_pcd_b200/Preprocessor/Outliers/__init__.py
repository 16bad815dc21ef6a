from .sor import SOR
