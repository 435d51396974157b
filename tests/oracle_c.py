"""The C oracle driver lives in oracle/c_driver.py; re-exported for the tests."""
from oracle.c_driver import COLL, ONT, PCG_DTYPE, COracle, OrcCfg, load  # noqa: F401
