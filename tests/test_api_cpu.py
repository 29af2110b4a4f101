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


def test_exact_pam_with_many_distinct_pams(oracle_engine):
    """256 distinct PAMs ('NNNN') do not fit the int8 categories the session returns: find_targets falls back to the code
    column; either way exact_pam equals a Categorical of the literal PAM strings (sorted categories)"""
    import numpy as np
    import pandas as pd
    from guidemaker_b200 import PamTarget

    class Rec:
        def __init__(self, i, s):
            self.id, self.seq = i, s

    rng = np.random.default_rng(17)
    recs = [Rec("r%d" % i, "".join(rng.choice(list("ACGT"), size=4000))) for i in range(2)]
    for pam, n_cat in (("NNNN", 256), ("NNG", 16)):
        df = PamTarget(pam, "3prime", "hamming").find_targets(recs, 12)
        assert len(df["exact_pam"].cat.categories) == n_cat
        assert list(df["exact_pam"].cat.categories) == sorted(df["exact_pam"].cat.categories)
        # forward rows: the PAM is the text right behind the target
        fwd = df[df["strand"] & (df["seqid"] == "r0")]
        lit = [recs[0].seq[int(e): int(e) + len(pam)] for e in fwd["stop"]]
        assert fwd["exact_pam"].astype(str).tolist() == lit
        want = pd.Categorical(df["exact_pam"].astype(str))
        assert np.array_equal(df["exact_pam"].cat.codes, want.codes)
