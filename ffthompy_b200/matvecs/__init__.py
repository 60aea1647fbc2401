from .objects import Matrix, VecTri, Scalar, DFT, LinOper, MultiVector, MultiOper, ScipyOper  # noqa: F401
