"""GPU tests added after the round's last GPU run (the builder's GPU budget was spent): kept in the file pytest collects LAST so that
`-x` reaches them only after every test that has already been seen green on a B200.  They cover `dataops.voxelize`, the workspace
forms of the op-level kNN entry points and `engine.evaluate_on_batches`; each is built from entry points the earlier files test."""
import pytest
import torch

from flowcompare_b200 import dataops, engine as eng, lib as fclib, spec
from oracle import dataops_ref, knn_ref
from tests.conftest import load_golden, voxel_label_mismatch
from oracle.make_golden import fixture_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def lib():
    return fclib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _engine(name, precision="fp32"):
    cfg, fsd, esd, batch = fixture_inputs(name)
    return cfg, fsd, esd, batch, eng.FlowCompareB200((fsd, esd), cfg, device=DEV, precision=precision)


@pytest.mark.parametrize("case", ["loader_default", "fine", "single_layer"])
def test_voxelize_matches_oracle_and_reference(case):
    """`voxelize` (reference utils.py:446-454): centres bit-exact against the UNMODIFIED reference (tests/golden/voxelize.pt), labels
    bit-exact against the canonical kNN oracle and equal to the reference's except at near-ties inside its own rounding noise."""
    from oracle.make_voxelize_golden import inputs as vox_inputs
    gold = load_golden("voxelize")[case]
    pos, start, end, size = vox_inputs(case)
    labels, centers = dataops.voxelize(pos.cuda(), start, end, size)
    assert labels.dtype == torch.int64 and tuple(labels.shape) == (pos.shape[0], 1) and centers.is_cuda
    assert torch.equal(centers.cpu(), gold["centers"])
    want, _ = dataops_ref.voxelize(pos, start, end, size)
    assert torch.equal(labels.cpu(), want)
    assert voxel_label_mismatch(pos, gold["centers"], labels.cpu(), gold["labels"]) <= 5e-3


def test_knn_workspace_forms_equal_the_scratch_forms(lib):
    """fc_knn_self_ws / fc_knn_query_ws (caller-provided workspace, no global state) give the indices of fc_knn_self / fc_knn_query,
    also from a misaligned workspace base and when two calls run on different streams with their own workspaces."""
    g = torch.Generator().manual_seed(11)
    pts = torch.randn(2, 700, 64, generator=g).to(DEV)
    want = eng.knn(pts.permute(0, 2, 1), 40)
    assert torch.equal(want.cpu(), knn_ref.knn_self(pts.cpu(), 40))
    nbytes = lib.fc_knn_workspace_bytes(2, 700, 700, 1)
    assert nbytes > 0
    ws = [torch.empty(nbytes + 8, dtype=torch.uint8, device=DEV) for _ in range(2)]
    got = [torch.full((2, 700, 40), -7, dtype=torch.int64, device=DEV) for _ in range(2)]
    streams = [torch.cuda.Stream(device=DEV) for _ in range(2)]
    torch.cuda.synchronize()
    for i in range(2):
        with torch.cuda.stream(streams[i]):
            rc = lib.fc_knn_self_ws(pts.data_ptr(), 64, 2, 700, 64, 40, 0, got[i].data_ptr(), ws[i].data_ptr() + 8 * i, nbytes,
                                    streams[i].cuda_stream)
            fclib.check(rc, "fc_knn_self_ws")
    torch.cuda.synchronize()
    assert torch.equal(got[0], want) and torch.equal(got[1], want)
    q, t = torch.randn(900, 3, generator=g).to(DEV), torch.randn(315, 3, generator=g).to(DEV)
    want_q = eng.get_knn(q, t, 1)
    nb = lib.fc_knn_workspace_bytes(1, 900, 315, 0)
    wq = torch.empty(nb, dtype=torch.uint8, device=DEV)
    got_q = torch.empty(900, 1, dtype=torch.int64, device=DEV)
    assert lib.fc_knn_query_ws(q.data_ptr(), t.data_ptr(), 900, 315, 3, 1, got_q.data_ptr(), wq.data_ptr(), nb, _stream()) == 0
    torch.cuda.synchronize()
    assert torch.equal(got_q, want_q) and torch.equal(got_q.cpu(), knn_ref.knn_query(q.cpu(), t.cpu(), 1))
    assert lib.fc_knn_query_ws(q.data_ptr(), t.data_ptr(), 900, 315, 3, 1, got_q.data_ptr(), wq.data_ptr(), nb - 512, _stream()) == -4




def test_evaluate_on_batches_equals_the_reference_loop_composed_by_hand():
    """engine.evaluate_on_batches = the loop of evaluate_on_test (reference test_flow.py:147-226): per batch the 1|0 and 0|0
    passes, log_prob_to_change, per-cloud change means, running average of the 1|0 pass's nats -- here with the two passes
    stacked into one, so it must equal the same loop written with separate calls (a pair's result does not depend on its batch)."""
    torch.set_grad_enabled(False)
    cfg, fsd, esd, batch, e = _engine("tiny_dgcnn_attn_extra", "tf32x3")
    B = batch["extract_0"].shape[0]
    data, epss = [], []
    for i in range(3):
        pair, ee = [], []
        for j in range(2):
            b = spec.synthetic_batch(cfg, B, seed=60 + 2 * i + j)
            pair.append((b["extract_0"].to(DEV), b["extract_1"].to(DEV), b["extra_context"].to(DEV)))
            ee.append(b["eps"].to(DEV))
        data.append(tuple(pair))
        epss.append(torch.cat(ee, dim=0))
    nats_avg, change_means = e.evaluate_on_batches(data, multiple=1.0, eps=epss)
    want_nats, want_means = 0.0, []
    for i, (b10, b00) in enumerate(data):
        _, lp10, nats = e.inner_loop(b10, eps=epss[i][:B])
        _, lp00, _ = e.inner_loop(b00, eps=epss[i][B:])
        change = eng.log_prob_to_change(lp10, lp00, 1.0)
        want_means.extend((change > 0).float().mean(dim=-1).tolist())
        want_nats = (want_nats * i + nats.item()) / (i + 1)
    assert isinstance(nats_avg, float) and len(change_means) == 3 * B
    assert change_means == want_means
    assert abs(nats_avg - want_nats) <= 1e-5 * max(1.0, abs(want_nats))
    e.close()


@pytest.mark.parametrize("n,k,sites", [(1500, 40, 1), (2000, 40, 3), (3000, 64, 2)])
def test_knn_masses_of_duplicate_points_follow_the_tie_rule(n, k, sites):
    """Hundreds of IDENTICAL points (the dataset's oversampling produces exact duplicates, reference utils.py:362-370): more than
    256 keys tie at or above a query's k-th best, which takes knn_select_kernel's exact-bisection path (ties in index order,
    across segments and the carried list).  Bit-exact against the canonical oracle; the algorithm itself is also checked on CPU
    by a lane-by-lane model (tests/knn_select_sim.py)."""
    g = torch.Generator().manual_seed(n + k)
    x = torch.rand(1, n, 6, generator=g)
    site = torch.rand(sites, 6, generator=g)
    pick = torch.rand(n, generator=g) < 0.6                      # 60 % of the points sit exactly on one of the sites
    x[0, pick] = site[torch.randint(0, sites, (int(pick.sum()),), generator=g)]
    want = knn_ref.knn_self(x, k)
    got = eng.knn(x.permute(0, 2, 1).to(DEV), k).cpu()
    assert torch.equal(got, want)
