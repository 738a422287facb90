"""Pins the CPU oracle (oracle/lin_reg_oracle.py) against the reference's own known-answer tests (tests/cases.py).

CPU only.  This is what makes the oracle trustworthy as the checker of the CUDA path: same seeds / literals, same
external libraries and tolerances the reference's test-suite uses on the reference implementation.
"""
import pytest

from tests import cases
from tests.backends import OracleBackend

BE = OracleBackend()


@pytest.mark.parametrize("case", cases.ALL_CASES, ids=lambda f: f.__name__)
def test_oracle_f64(case, monkeypatch):
    import polars_ds_extension_b200.config as cfg

    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", True)
    case(BE)


@pytest.mark.parametrize("case", cases.ALL_CASES, ids=lambda f: f.__name__)
def test_oracle_f32(case, monkeypatch):
    import polars_ds_extension_b200.config as cfg

    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", False)
    case(BE)
