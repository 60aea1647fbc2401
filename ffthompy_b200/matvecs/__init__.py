from .objects import Matrix, VecTri, Scalar, DFT, LinOper  # noqa: F401
