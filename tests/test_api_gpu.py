"""The drop-in API on the real CUDA engine (libgm_b200.so through ctypes): the reference-style cases of
tests/api_cases.py against the reference fixtures, bit-exact.
"""
import pytest

from tests import api_cases as C

pytestmark = pytest.mark.gpu


def test_pam_attributes(cuda_engine):
    C.case_pam_attributes()


def test_find_targets_inline(cuda_engine, inline_ref):
    C.case_find_targets_inline(inline_ref)


@pytest.mark.parametrize("name", list(C.CARSONELLA_CASES))
def test_carsonella(cuda_engine, name, carsonella, carsonella_ref, config_yaml):
    C.case_carsonella(name, carsonella, carsonella_ref, config_yaml)


def test_handmade_frame(cuda_engine, config_yaml):
    C.case_handmade_frame(config_yaml)


def test_levin_dist(cuda_engine, inline_ref, config_yaml):
    C.case_levin_dist(inline_ref, config_yaml)


@pytest.mark.parametrize("name", C.SYNTH)
def test_synthetic(cuda_engine, name, synthetic_ref, config_yaml):
    C.case_synthetic(name, synthetic_ref, config_yaml)


def test_controls(cuda_engine, controls_ref, carsonella, synthetic_ref, config_yaml, tmp_path):
    C.case_controls(controls_ref, carsonella, synthetic_ref, config_yaml, tmp_path)


def test_errors(cuda_engine, config_yaml):
    C.case_errors(config_yaml)


@pytest.mark.gpu
def test_unsorted_contigs(cuda_engine, config_yaml):
    C.case_unsorted_contigs(config_yaml)
