"""Host graph tables vs the reference's getGraphDict (when present) and vs goldens."""
import numpy as np
import pytest
import torch

from sgrl_b200 import graph as G, morphologies as M
from oracle import ref_loader
import parity


def test_morphology_inventory():
    assert len(M.CWHH) == 23 and len(M.ALL) == 27
    for name, par in M.ALL.items():
        assert par[0] == -1 and len(par) <= M.MAX_LIMBS
        assert int(name.split("_")[2]) == len(par)
        assert all(0 <= p < i for i, p in enumerate(par) if i > 0)


def test_hopper5_ranks():
    t = G.traversal_ranks([-1, 0, 1, 2, 3])
    assert t == [[0, 1, 2, 3, 4], [4, 3, 2, 1, 0], [4, 3, 2, 1, 0]]


def test_relation_against_golden():
    gold = parity.load_golden()
    for name, _ in parity.CASES:
        g = G.build_graph(M.ALL[name])
        np.testing.assert_array_equal(torch.stack(g["traversals"]).numpy(), gold[name + "/traversals"])
        np.testing.assert_allclose(g["relation"].numpy(), gold[name + "/relation"], rtol=0, atol=2e-6)
        assert g["relation"].shape == (len(M.ALL[name]),) * 2 + (3,)


@pytest.mark.skipif(ref_loader.find_reference() is None, reason="reference not on this machine")
def test_graph_against_reference():
    ref = ref_loader.load_reference()
    for name, par in M.ALL.items():
        want = ref.utils.getGraphDict(par, ["pre", "inlcrs", "postlcrs"], device=torch.device("cpu"))
        got = G.build_graph(par)
        for a, b in zip(want["traversals"], got["traversals"]):
            assert torch.equal(a, b), name
        assert torch.equal(want["relation"], got["relation"]), name
    assert G.build_graph([-1]) == {"parents": [-1]}
