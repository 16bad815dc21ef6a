"""Scaler blocks.  A scaler is a preprocessing block that also remembers the affine map it applied
(``mean`` (1,3) and ``scale``), which ``Aligner.transfrom`` reads back to place a raw cloud in the
registered frame (reference: Aligner/Aligner.py:373-380).  Host numpy code: it fixes every downstream
input bit for bit and costs O(N)."""
from .BaseScaler import BaseScaler
from .radiusScaler import RadiusScaler
