"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/fmd_b200.h declares (and nothing is declared only in Python), argument validation works
without a GPU, and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fmd_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fmd_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from flashmd import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes table binds exactly the declared compute entry points
    bound = set(_lib.EXPORTED)
    assert bound == set(names), (sorted(bound - set(names)), sorted(set(names) - bound))


def test_every_entry_point_cites_the_reference_interface_it_replaces():
    src = open(HEADER).read()
    blocks = re.findall(r"/\*(.*?)\*/\s*int\s+(fmd_[a-z0-9_]+)\s*\(", src, flags=re.S)
    assert len(blocks) >= 25
    uncited = [n for c, n in blocks if "replaces" not in c and "contract as" not in c and "Same physics" not in c
               and n not in ("fmd_version", "fmd_sm_count", "fmd_exclusive_scan_i32", "fmd_increment_u64",
                             "fmd_philox_normal", "fmd_nl_fill", "fmd_nl_reverse", "fmd_debug_set_trace_fwd",
                             "fmd_debug_set_trace_bwd")]
    assert not uncited, uncited


def test_argument_validation_without_gpu():
    from flashmd import _lib as L
    lib = L.load()
    assert lib.fmd_version() >= 100
    rc = lib.fmd_linear(None, 0, None, 0, None, None, 0, 4, 4, 4, None, 0, 0, 0, None, 0, None, None)
    assert rc == -1 and b"fmd_linear" in lib.fmd_last_error()
    rc = lib.fmd_prior_energy_forces(9, None, None, None, 0, None, None, None, 1, None, None, None)
    assert rc == -1 and b"unknown prior kind" in lib.fmd_last_error()


def test_no_cpu_fallback():
    from flashmd import _lib as L
    from flashmd.engine import ForceField
    with pytest.raises(RuntimeError):
        L.ptr(torch.zeros(3))
    with pytest.raises(RuntimeError):
        ForceField(None, [], torch.zeros(4, dtype=torch.long), torch.tensor([0, 4]))


def test_prior_incidence_lists_packed_records_decode_to_the_inputs():
    """Host logic of the owner-computes prior layout (engine.PriorCSR): every pair term is listed under both of its
    beads, records are 8 bytes {other | kind << 28, id} and the deduplicated parameter table returns exactly the
    parameters of the term (bonds k, x0, V0; repulsion sigma)."""
    from flashmd import _lib as L
    from flashmd.engine import PriorCSR, PriorTerm
    g = torch.Generator().manual_seed(0)
    n_nodes = 40
    bi = torch.stack([torch.arange(0, n_nodes - 1), torch.arange(1, n_nodes)]).to(torch.int32)
    kinds = torch.randint(0, 3, (bi.shape[1],), generator=g)
    k = torch.tensor([10.0, 20.0, 30.0])[kinds]
    x0 = torch.tensor([3.8, 4.0, 4.2])[kinds]
    v0 = torch.zeros_like(k)
    ri, rj = torch.triu_indices(n_nodes, n_nodes, offset=2)
    rep = torch.stack([ri, rj]).to(torch.int32)
    sigma = torch.tensor([3.0, 3.5])[torch.randint(0, 2, (rep.shape[1],), generator=g)]
    priors = [PriorTerm(L.PRIOR_BONDS, bi, torch.zeros(bi.shape[1], dtype=torch.int32), k, x0, v0),
              PriorTerm(L.PRIOR_REPULSION, rep, torch.zeros(rep.shape[1], dtype=torch.int32), sigma)]
    csr = PriorCSR(priors, n_nodes, "cpu")
    assert csr.pair_tab is not None and csr.pair_ent.shape[1] == 2 and csr.pair_tab.shape[0] <= 3 + 2
    ptr, ent, tab = csr.pair_ptr.long(), csr.pair_ent, csr.pair_tab
    assert int(ptr[-1]) == 2 * (bi.shape[1] + rep.shape[1])
    want = {}
    for (m, kind, ps) in ((bi, L.PRIOR_BONDS, (k, x0, v0)), (rep, L.PRIOR_REPULSION, (sigma, torch.zeros_like(sigma), torch.zeros_like(sigma)))):
        for t in range(m.shape[1]):
            a, b = int(m[0, t]), int(m[1, t])
            want[(a, b)] = want[(b, a)] = (kind, tuple(float(p[t]) for p in ps))
    seen = 0
    for a in range(n_nodes):
        for p in range(int(ptr[a]), int(ptr[a + 1])):
            head, pid = int(ent[p, 0]), int(ent[p, 1])
            other, kind = head & 0x0FFFFFFF, head >> 28
            kw, pw = want[(a, other)]
            assert kind == kw and tuple(float(v) for v in tab[pid, :3]) == pw
            seen += 1
    assert seen == len(want)
