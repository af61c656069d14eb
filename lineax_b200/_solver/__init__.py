from .bicgstab import BiCGStab as BiCGStab
from .cg import CG as CG, NormalCG as NormalCG
from .cholesky import Cholesky as Cholesky
from .diagonal import Diagonal as Diagonal
from .gmres import GMRES as GMRES
from .lsmr import LSMR as LSMR
from .lu import LU as LU
from .normal import Normal as Normal
from .qr import QR as QR
from .triangular import Triangular as Triangular
from .tridiagonal import Tridiagonal as Tridiagonal
