"""Type aliases of the reference (python/polars_ds/typing.py) that the lin_reg family uses."""
from typing import Literal

LRSolverMethods = Literal["qr", "svd", "cholesky", "choleskey"]
NullPolicy = Literal["raise", "skip", "zero", "one", "ignore"]
