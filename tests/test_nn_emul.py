"""CPU check of the device NN search LOGIC: nn_core.cuh (the function every thread of k_nn_search runs) compiled for
the host by tests/csrc/nn_emul.cpp, against the oracle's restatement of kernel_semanticNearestNeighborSearch
(lesson_16.cu:531-703) on the same cases the GPU parity tests use.  Bit-exact indices required."""
import numpy as np
import pytest

from tests import native
from tests.test_gpu_stages import _cases


def _check(oracle, first, second, radius, ext, max_in, max_out, name):
    nn_o, gp, table, buckets = oracle.semantic_nn(first, second, radius, radius, ext, max_in, max_out)
    nn_e, ev = native.nn_emul_search(first, second, table, buckets, gp, radius, max_in, max_out, prune=True)
    assert np.array_equal(nn_e, nn_o), (name, int((nn_e != nn_o).sum()))
    nn_u, ev_u = native.nn_emul_search(first, second, table, buckets, gp, radius, max_in, max_out, prune=False)
    assert np.array_equal(nn_u, nn_o), (name, "unpruned")
    assert ev <= ev_u
    return ev, ev_u


def test_emulated_search_matches_oracle(synth, oracle):
    for name, first, second, radius, ext, max_in, max_out in _cases(synth):
        if len(first) > 120000:
            first, second = first[:120000], second[:8000]
        _check(oracle, first, second, radius, ext, max_in, max_out, name)


def test_emulated_search_scan_pairs(synth, oracle):
    """Scan-like inputs (surfaces, coherent labels) incl. radius != bucket size and a perturbed pair."""
    first, second, pose_init, pose2, _ = synth.scan_pair("hdl32", seed=11, n_azimuth=512)
    fg = oracle.transform_cloud(first, pose_init)
    sg = oracle.transform_cloud(second, pose2)
    ev, ev_u = _check(oracle, fg, sg, 0.5, 1.0, 100, 100, "hdl_pair")
    assert ev * 4 < ev_u          # the rounds prune most of the neighbourhood
    rng = np.random.default_rng(0)
    sg2 = sg[rng.permutation(len(sg))].copy()
    for radius, bucket in ((2.5, 2.5), (0.3, 1.0), (1.0, 0.4)):
        nn_o, gp, table, buckets = oracle.semantic_nn(fg, sg2, radius, bucket)
        nn_e, _ = native.nn_emul_search(fg, sg2, table, buckets, gp, radius)
        assert np.array_equal(nn_e, nn_o), (radius, bucket)
