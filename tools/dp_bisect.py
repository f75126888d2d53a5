"""Single-GPU bisection of the data-parallel C3 step (loss differs between world sizes).

Part A - shapes: the gradient of the full 8192-row batch must equal the SUM of the gradients of its N row shards
         (what the NCCL all-reduce computes) to fp64 re-association accuracy, for N = 2, 4, 8.  A shape-specific GEMM
         defect at a 2048-row shard shows up here without any second GPU.
Part B - schedules: the same 9 Adam steps with every combination of {optimizer overlapped on the side stream | after
         backward} x {launch tape | planner path}, repeated; all must agree BITWISE (same kernels, same operands).
Under torchrun (world > 1) part B runs with the NCCL all-reduce and compares every schedule with the plain one
(synchronous all-reduce, no overlap, no tape) and across repetitions."""
import os, sys, time, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad
from autodiff_b200 import dp
from autodiff_b200.core import tape, engine

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
ad.init(local)
dist = None
if world > 1 or "nccl1" in sys.argv[1:]:
    import torch.distributed as dist
    if world == 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

WIDTH = int(os.environ.get("BISECT_WIDTH", "4096"))
LAYERS = int(os.environ.get("BISECT_LAYERS", "8"))
GB = int(os.environ.get("BISECT_BATCH", "8192"))
STEPS = int(os.environ.get("BISECT_STEPS", "9"))
sizes = [WIDTH] * (LAYERS + 1)
np.random.seed(1337)
W0 = [np.random.randn(a + 1, b) * 0.01 for a, b in zip(sizes, sizes[1:])]
rng = np.random.default_rng(1337)
X = rng.standard_normal((GB, sizes[0]))
lab = rng.integers(0, sizes[-1], GB)
Y = np.zeros((GB, sizes[-1])); Y[np.arange(GB), lab] = 1.0


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def fresh_nn():
    np.random.seed(0)
    nn = ad.NN([2, 2])
    nn.sizes = sizes
    nn.w = [ad.Variable(ad.to_device(w), name="w%d" % k) for k, w in enumerate(W0)]
    return nn


def grads_of(nn, xd, yd):
    xv, yv = ad.Variable(xd, name="images"), ad.Variable(yd, name="labels")
    cost = ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=yv, logits=nn(xv))) / GB
    gs = ad.grad(cost, nn.w)
    ad.evaluate_many(list(reversed(gs)) + [cost])
    return [ad.to_host(g.eval_device()) for g in gs], float(ad.to_host(cost.eval_device()))


def part_a():
    nn = fresh_nn()
    tape.ENABLED = False
    full, loss_full = grads_of(nn, ad.to_device(X), ad.to_device(Y))
    scale = [float(np.max(np.abs(g))) for g in full]
    for n in (2, 4, 8):
        for use_tape in (False, True):
            tape.ENABLED = use_tape
            tape.clear()
            acc, loss = None, 0.0
            reps = 3 if use_tape else 1           # 3rd sighting of a structure = first replay
            for r in range(n):
                lo, hi = dp.shard_rows(GB, n, r)
                xd, yd = ad.to_device(X[lo:hi]), ad.to_device(Y[lo:hi])
                for _ in range(reps):
                    g, l = grads_of(nn, xd, yd)
                acc = g if acc is None else [a + b for a, b in zip(acc, g)]
                loss += l
            errs = [float(np.max(np.abs(a - f))) / s for a, f, s in zip(acc, full, scale)]
            say("A: N=%d rows=%d tape=%s: loss rel err %.2e, grad rel err per layer %s  %s"
                % (n, GB // n, use_tape, abs(loss - loss_full) / abs(loss_full), " ".join("%.1e" % e for e in errs),
                   "OK" if max(errs) < 1e-12 else "<<<<<<<<<< MISMATCH"))
    tape.ENABLED = True
    tape.clear()


def train(rows, overlap, use_tape, d, steps=STEPS, sync_allreduce=False):
    tape.ENABLED = use_tape
    tape.clear()
    nn = fresh_nn()
    opt = ad.Adam(len(nn.w), lr=1e-3)
    lo, hi = rows
    xd, yd = ad.to_device(X[lo:hi]), ad.to_device(Y[lo:hi])
    losses = []
    if sync_allreduce:
        saved = dp.allreduce_async
        class _Done:
            def wait(self):
                pass
        def sync_ar(arr, dd):
            dd.all_reduce(dp._flat_view(arr), op=dd.ReduceOp.SUM)
            torch.cuda.synchronize()
            return _Done()
        dp.allreduce_async = sync_ar
    try:
        for _ in range(steps):
            losses.append(dp.train_step(nn, opt, xd, yd, GB, d, read_loss=True, overlap_optimizer=overlap))
    finally:
        if sync_allreduce:
            dp.allreduce_async = saved
    torch.cuda.synchronize()
    crc = 0
    for w in nn.w:
        crc = zlib.crc32(np.ascontiguousarray(w.value).tobytes(), crc)
    return losses, crc, dict(tape.stats)


def part_b():
    if world == 1:
        row_sets = [(0, GB), (0, GB // 4), (GB // 4, GB // 2), (0, GB // 8)]
        if os.environ.get("BISECT_QUICK"):
            row_sets = [(0, GB // 4)]
    else:
        row_sets = [dp.shard_rows(GB, world, rank)]
    reps = int(os.environ.get("BISECT_REPS", "3"))
    for rows in row_sets:
        base = None
        configs = [("plain", False, False, True)] if dist is not None else [("plain", False, False, False)]
        configs += [("tape", False, True, False), ("overlap", True, False, False), ("overlap+tape", True, True, False)]
        if os.environ.get("BISECT_QUICK"):
            configs = [c for c in configs if c[0] in ("plain", "overlap")]
        for name, overlap, use_tape, sync_ar in configs:
            for rep in range(reps if name != "plain" else 2):
                t0 = time.perf_counter()
                losses, crc, st = train(rows, overlap, use_tape, dist, sync_allreduce=sync_ar)
                if base is None:
                    base = (losses, crc)
                same = losses == base[0] and crc == base[1]
                if dist is not None and world > 1:       # every rank must hold the same weights
                    t = torch.tensor([crc], dtype=torch.int64, device="cuda")
                    lst = [torch.zeros_like(t) for _ in range(world)]
                    dist.all_gather(lst, t)
                    ranks_same = len({int(v.item()) for v in lst}) == 1
                else:
                    ranks_same = True
                say("B: world=%d rows=%s %-13s rep %d: last loss %.15g  crc %08x  %s%s  (%.1f s, replays %d)"
                    % (world, rows, name, rep, losses[-1], crc, "bitwise == plain" if same else "<<<<<<<<<< DIFFERS: first bad step %s"
                       % next((i for i, (a, b) in enumerate(zip(losses, base[0])) if a != b), "weights only"),
                       "" if ranks_same else "  RANKS DISAGREE", time.perf_counter() - t0, st["replays"]))


def part_c():
    """Which array goes wrong first?  Plain vs overlapped schedule, state compared after every step."""
    rows = (0, GB // 4)
    def run(overlap, nsteps=3):
        tape.ENABLED = False
        nn = fresh_nn()
        opt = ad.Adam(len(nn.w), lr=1e-3)
        xd, yd = ad.to_device(X[rows[0]:rows[1]]), ad.to_device(Y[rows[0]:rows[1]])
        snaps = []
        for _ in range(nsteps):
            dp.train_step(nn, opt, xd, yd, GB, None, read_loss=True, overlap_optimizer=overlap)
            torch.cuda.synchronize()
            snaps.append([(np.array(w.value), ad.to_host(opt.m[k]), ad.to_host(opt.v[k])) for k, w in enumerate(nn.w)])
        return snaps
    a = run(False)
    for rep in range(2):
        b = run(True)
        for s, (sa, sb) in enumerate(zip(a, b)):
            for k, (ta, tb) in enumerate(zip(sa, sb)):
                for name, x, y in zip("wmv", ta, tb):
                    bad = np.argwhere(x != y)
                    if len(bad):
                        r, c = bad[:, 0], bad[:, 1]
                        say("C rep %d step %d layer %d %s: %d of %d elements differ; rows %d..%d (%d distinct) cols %d..%d (%d distinct); max |diff| %.3e, max |ref| %.3e"
                            % (rep, s, k, name, len(bad), x.size, r.min(), r.max(), len(np.unique(r)), c.min(), c.max(), len(np.unique(c)),
                               float(np.max(np.abs(x - y))), float(np.max(np.abs(x)))))
            if any((x != y).any() for ta, tb in zip(sa, sb) for x, y in zip(ta, tb)):
                break


def part_d():
    """First corrupted NODE: every node value of step 1 (second step), plain vs overlapped, in graph order."""
    rows = (0, GB // 4)
    def walk(tops):
        seen, out, stack = set(), [], list(reversed(tops))
        while stack:
            n = stack.pop()
            if id(n) in seen:
                continue
            seen.add(id(n))
            out.append(n)
            stack.extend(reversed(n.children))
        return out
    def run(overlap):
        tape.ENABLED = False
        nn = fresh_nn()
        opt = ad.Adam(len(nn.w), lr=1e-3)
        xd, yd = ad.to_device(X[rows[0]:rows[1]]), ad.to_device(Y[rows[0]:rows[1]])
        dp.train_step(nn, opt, xd, yd, GB, None, read_loss=True, overlap_optimizer=overlap)
        keep = []
        dp.train_step(nn, opt, xd, yd, GB, None, read_loss=True, overlap_optimizer=overlap)
        torch.cuda.synchronize()
        vals = []
        order = walk(keep)
        pos = {id(n): i for i, n in enumerate(order)}
        for n in order:
            d = n._dev
            vals.append((type(n).__name__ + ("[" + n.name[:12] + "]" if type(n).__name__ == "Variable" else ""), getattr(n, "op_str", ""), n.shape,
                         None if d is None or not hasattr(d, "shape") else ad.to_host(d), [pos[id(c)] for c in n.children]))
        return vals
    a = run(False)
    for rep in range(int(os.environ.get('BISECT_REPS', '2'))):
        b = run(True)
        nbad = 0
        for i, (na, nb) in enumerate(zip(a, b)):
            x, y = na[3], nb[3]
            if x is None or y is None or x.shape != y.shape:
                continue
            if not np.array_equal(x, y, equal_nan=True):
                bad = np.argwhere(np.atleast_2d(x) != np.atleast_2d(y))
                r, c = bad[:, 0], bad[:, -1]
                say("D rep %d node %d %s %s %s: %d of %d differ; rows %d..%d (%d distinct), cols %d..%d (%d distinct); max|diff| %.3e max|ref| %.3e; children %s first %s"
                    % (rep, i, na[0], na[1], na[2], len(bad), x.size, r.min(), r.max(), len(np.unique(r)), c.min(), c.max(), len(np.unique(c)),
                       float(np.max(np.abs(x - y))), float(np.max(np.abs(x))), na[4], bad[:4].tolist()))
                nbad += 1
                if nbad >= int(os.environ.get("BISECT_MAXBAD", "80")):
                    break
        say("D rep %d: %d nodes compared, %d differ" % (rep, len(a), nbad))


if "c" in sys.argv[1:]:
    part_c()
if "d" in sys.argv[1:]:
    part_d()

if "a" in sys.argv[1:] or not [a for a in sys.argv[1:] if a in ("a", "b", "c", "d")]:
    if world == 1:
        part_a()
if "b" in sys.argv[1:] or not [a for a in sys.argv[1:] if a in ("a", "b", "c", "d")]:
    part_b()
if dist is not None:
    dist.destroy_process_group()
