"""thallo_b200: B200-native (sm_100a) Gauss-Newton / Levenberg-Marquardt + PCG solver
backend behind Thallo's C ABI.  See DESIGN.md."""
__version__ = "0.1.0"
