from .cg import CG as CG, NormalCG as NormalCG
from .lu import LU as LU
