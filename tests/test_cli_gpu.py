"""The CLI runner on the real CUDA engine: same checks as tests/test_cli_cpu.py."""
import pytest

from tests import test_cli_cpu as T

pytestmark = pytest.mark.gpu


def test_cli_raw_output_and_offtargets_gpu(cuda_engine, tmp_path, carsonella_ref):
    T.test_cli_raw_output_and_offtargets(cuda_engine, tmp_path, carsonella_ref)
