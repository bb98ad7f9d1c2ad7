"""CPU check of the device NN search LOGIC: nn_core.cuh (the function every thread of k_nn_search runs) compiled for
the host by tests/csrc/nn_emul.cpp, against the oracle's restatement of kernel_semanticNearestNeighborSearch
(lesson_16.cu:531-703) on the same cases the GPU parity tests use.  Bit-exact indices required."""
import numpy as np
import pytest

from tests import native
from tests.test_gpu_stages import _cases, _spatial_order


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


def test_warp_shared_algorithm_matches_oracle(synth, oracle):
    """The ALGORITHM of k_nn_search_hull / k_nn_search_grid (hull of the lanes' boxes, representative cells, with and without previous-hull skipping, staged
    groups with group minimum / tie flag / cold step / re-scan, hull-based settle test, scattered-warp fallback), restated
    for the host in tests/csrc/nn_emul.cpp, against the oracle: sorted and unsorted queries, pruning on and off, radius
    above and below the bucket size, aliasing labels, bins larger than a staging batch, exact ties."""
    cases = []
    for name, first, second, radius, ext, max_in, max_out in _cases(synth):
        if max_in != max_out or max_out <= 0 or len(first) < 2:
            continue                              # two candidate sets use the per-thread kernel
        if len(first) > 60000:
            first, second = first[:60000], second[:6000]
        cases.append((name, first, second[:8000], radius, radius, max_out))
    f, q = synth.random_cloud(40000, seed=61, extent=(3, 2, 0.05)), synth.random_cloud(6000, seed=62, extent=(3, 2, 0.05))
    cases += [("r_gt_b", f, q, 1.0, 0.4, 100), ("r_lt_b", f, q, 0.3, 1.0, 100), ("r_2b", f, q, 2.0, 1.0, 100)]
    fa, qa = f.copy(), q.copy()
    fa["label"] = (fa["label"] % 3) * 4
    qa["label"] = (qa["label"] % 3) * 4
    cases += [("alias_labels", fa, qa, 0.5, 0.5, 100)]
    cases += [("one_bin", synth.random_cloud(20000, seed=63, extent=(0.05, 0.05, 0.05), n_labels=1),
               synth.random_cloud(2000, seed=64, extent=(0.3, 0.3, 0.3), n_labels=1), 1.0, 1.0, 1000)]
    sf, ss, p1, p2, _ = synth.scan_pair("sick", seed=65, n_beams=128, n_profiles=128)
    cases += [("sick_scan", oracle.transform_cloud(sf, p1), oracle.transform_cloud(ss, p2), 1.0, 1.0, 100)]
    shared_total, rescans_total = 0, 0
    for name, first, second, radius, bucket, cap in cases:
        for order in ("sorted", "mixed_labels", "as_is"):
            sq = second if order == "as_is" else second[_spatial_order(second, by_label=(order == "sorted"))].copy()
            nn_o, gp, table, buckets = oracle.semantic_nn(first, sq, radius, bucket, 1.0, cap, cap)
            for prune, skip_old in ((True, False), (False, False), (True, True)):      # k_nn_search_hull; round 1's hull skipping
                nn_w, fb, rs = native.nn_emul_search_warp(first, sq, table, buckets, gp, radius, cap, prune, skip_old_hull=skip_old)
                assert np.array_equal(nn_w, nn_o), (name, order, prune, skip_old, int((nn_w != nn_o).sum()))
                if prune and not skip_old:
                    shared_total += len(sq) - fb
                    rescans_total += rs
    assert shared_total > 20000           # the warp-shared path (not the fallback) answered a good part of the queries
    assert rescans_total > 0              # and the tie / inadmissible-winner re-scan was exercised (lattice ties, random normals)
