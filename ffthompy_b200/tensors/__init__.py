from .objects import Tensor, Scalar, einsum, scalar_product, norm_fun
from .operators import (DFT, grad, div, laplace, symgrad, potential, Operator, matrix2tensor, vector2tensor,
                        grad_div_tensor, grad_tensor, div_tensor, outer)
