"""Host logic of the drop-in API (frame assembly, row order, masks, thresholds, control loop) on CPU,
with the C-ABI calls routed to the oracle by the `oracle_engine` fixture.  The same cases run on the
real CUDA engine in tests/test_api_gpu.py."""
import pytest

from tests import api_cases as C


def test_pam_attributes():
    C.case_pam_attributes()


def test_find_targets_inline(oracle_engine, inline_ref):
    C.case_find_targets_inline(inline_ref)


@pytest.mark.parametrize("name", list(C.CARSONELLA_CASES))
def test_carsonella(oracle_engine, name, carsonella, carsonella_ref, config_yaml):
    C.case_carsonella(name, carsonella, carsonella_ref, config_yaml)


def test_handmade_frame(oracle_engine, config_yaml):
    C.case_handmade_frame(config_yaml)


def test_levin_dist(oracle_engine, inline_ref, config_yaml):
    C.case_levin_dist(inline_ref, config_yaml)


@pytest.mark.parametrize("name", C.SYNTH)
def test_synthetic(oracle_engine, name, synthetic_ref, config_yaml):
    C.case_synthetic(name, synthetic_ref, config_yaml)


def test_controls(oracle_engine, controls_ref, carsonella, synthetic_ref, config_yaml, tmp_path):
    C.case_controls(controls_ref, carsonella, synthetic_ref, config_yaml, tmp_path)


def test_errors(oracle_engine, config_yaml):
    C.case_errors(config_yaml)


def test_unsorted_contigs(oracle_engine, config_yaml):
    C.case_unsorted_contigs(config_yaml)
