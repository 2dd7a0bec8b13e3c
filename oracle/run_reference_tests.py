"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Runs the reference's own unittest suites (bases/test, solver/test, test/,
stability/test_rbc1d.py) against the UNMODIFIED reference Python with the C
restatements of oracle/fortran_kernels.c in place of the f2py modules
(oracle/shim.py).  This pins the restatements to every test the reference holds
for the path.  Works only where /root/reference exists.

    python -m oracle.run_reference_tests
"""
import sys

import pytest

from . import shim


def main():
    shim.load_reference()
    root = shim.REFERENCE_ROOT
    args = ["-q", "-p", "no:cacheprovider", "--import-mode=importlib", "-W", "ignore",
            root + "/pypde/bases/test", root + "/pypde/solver/test", root + "/pypde/test",
            root + "/pypde/stability/test_rbc1d.py",
            "--deselect", root + "/pypde/bases/test/test_fourier.py"]
    return pytest.main(args + sys.argv[1:])


if __name__ == "__main__":
    sys.exit(main())
