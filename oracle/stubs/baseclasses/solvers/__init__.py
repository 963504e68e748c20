from .. import BaseSolver  # noqa: F401  (reference FEMpy/Model.py:24 imports it from here)
