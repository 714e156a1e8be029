"""CPU tests of the host-side logic of the product: the boundary-condition stencil records that the
C-ABI layer ships to the device (host_logic.hpp) against the oracle's EIKONAL3D_SETBCS, and the
tile execution order of the CUDA kernel against the global hyperplane order (emulated on the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("shim") / "libshim.so"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           "-I", os.path.join(ROOT, "mceik_b200", "csrc"), "-I", "/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "host_logic_shim.cpp"), "-o", str(so)])
    return C.CDLL(str(so))


def _records(shim, nx, ny, nz, h, ts, xs, ys, zs, org=(0.0, 0.0, 0.0)):
    ts, xs, ys, zs = (np.ascontiguousarray(a, np.float64) for a in (ts, xs, ys, zs))
    cap = 27 * len(ts)
    node, col = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    d, t = np.zeros(cap), np.zeros(cap)
    p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    n = shim.shim_bc_records(nx, ny, nz, C.c_double(h), C.c_double(org[0]), C.c_double(org[1]), C.c_double(org[2]), len(ts),
                             p(ts, C.c_double), p(xs, C.c_double), p(ys, C.c_double), p(zs, C.c_double), cap,
                             p(node, C.c_int), p(d, C.c_double), p(t, C.c_double), p(col, C.c_int))
    return n, node[:max(n, 0)], d[:max(n, 0)], t[:max(n, 0)], col[:max(n, 0)]


def _oracle_bcs(nx, ny, nz, h, ts, xs, ys, zs, slow, org=(0.0, 0.0, 0.0)):
    n = nx * ny * nz
    lisbc = np.zeros(n, np.uint8)
    u = np.zeros(n)
    a = [np.ascontiguousarray(v, np.float64) for v in (ts, xs, ys, zs, slow)]
    p = lambda v: v.ctypes.data_as(O.c_dbl_p)
    ierr = O.lib().oracle_setbcs(nx, ny, nz, len(a[0]), C.c_double(h), C.c_double(h), C.c_double(h), C.c_double(org[0]),
                                 C.c_double(org[1]), C.c_double(org[2]), p(a[0]), p(a[1]), p(a[2]), p(a[3]), p(a[4]),
                                 lisbc.ctypes.data_as(C.POINTER(C.c_ubyte)), p(u))
    return ierr, lisbc, u


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_bc_records_reproduce_setbcs(shim, seed):
    rng = np.random.default_rng(seed)
    nx, ny, nz, h = 19, 23, 17, 75.0
    org = (100.0, -50.0, 12.5)
    slow = 1.0 / rng.uniform(2000, 6000, nx * ny * nz)
    ns = 3
    xs = org[0] + rng.uniform(h, (nx - 2) * h, ns)
    ys = org[1] + rng.uniform(h, (ny - 2) * h, ns)
    zs = org[2] + rng.uniform(h, (nz - 2) * h, ns)
    xs[1] = org[0] + 5 * h          # exactly on a node in x
    ys[2], zs[2] = ys[0], zs[0]     # overlapping stencils
    xs[2] = xs[0] + 0.25 * h
    ts = rng.uniform(0, 1, ns)
    n, node, d, t, col = _records(shim, nx, ny, nz, h, ts, xs, ys, zs, org)
    ierr, lisbc, u = _oracle_bcs(nx, ny, nz, h, ts, xs, ys, zs, slow, org)
    assert ierr == 0 and n > 0
    # replay the records the way apply_bcs_kernel does (sequential, min or assign)
    ur = np.full(nx * ny * nz, np.finfo(np.float64).max)
    for i in range(n):
        v = t[i] + d[i] * slow[node[i]]
        ur[node[i]] = v if col[i] else min(ur[node[i]], v)
    assert np.array_equal(ur, u)
    assert np.array_equal(np.unique(node), np.flatnonzero(lisbc))


def test_bc_records_error_cases(shim):
    nx = ny = nz = 10
    assert _records(shim, nx, ny, nz, 10.0, [0.0], [0.0], [45.0], [45.0])[0] == -1     # on node 1
    assert _records(shim, nx, ny, nz, 10.0, [0.0], [-4.0], [45.0], [45.0])[0] == -1    # left of the grid
    assert _records(shim, nx, ny, nz, 10.0, [0.0], [200.0], [45.0], [45.0])[0] == -1   # right of the grid
    n, node, d, t, col = _records(shim, nx, ny, nz, 10.0, [0.5], [80.0], [40.0], [40.0])  # node nx-1, on nodes in y,z
    assert n == 2 * 3 * 3 and col.sum() == 1 and d[col == 1][0] == 0.0
    assert _oracle_bcs(nx, ny, nz, 10.0, [0.0], [0.0], [45.0], [45.0], np.ones(1000))[0] == 1


@pytest.mark.parametrize("shape", [(37, 50, 21), (16, 16, 16), (33, 17, 48), (5, 70, 3), (64, 48, 32)])
def test_tile_order_equals_hyperplane_order(tmp_path, shape):
    """tests/tile_order_emulation.c: 16^3 tiles + clamped halo + tile-hyperplane order == global order, bit for bit."""
    O.build()
    exe = tmp_path / "tile_emu"
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-o", str(exe),
                           os.path.join(ROOT, "tests", "tile_order_emulation.c"), "-L", O.ORACLE_DIR, "-loracle", "-lm",
                           f"-Wl,-rpath,{O.ORACLE_DIR}"])
    out = subprocess.run([str(exe)] + [str(s) for s in shape], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("MATCH"), out.stdout


def _align(shim, eb, optr, tid, tobs, var):
    optr, tid = np.ascontiguousarray(optr, np.int32), np.ascontiguousarray(tid, np.int32)
    tobs, var = np.ascontiguousarray(tobs, np.float64), np.ascontiguousarray(var, np.float64)
    ne = optr.size - 1
    cap = max(1, 8 * 8 * max(1, tid.size))
    o2, t2 = np.zeros(ne + 1, np.int32), np.zeros(cap, np.int32)
    a2, b2 = np.zeros(cap), np.zeros(cap)
    p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    n = shim.shim_align_event_blocks(eb, ne, p(optr, C.c_int), p(tid, C.c_int), p(tobs, C.c_double), p(var, C.c_double), cap,
                                     p(o2, C.c_int), p(t2, C.c_int), p(a2, C.c_double), p(b2, C.c_double))
    assert n >= 0
    return o2, t2[:n], a2[:n], b2[:n]


def test_align_event_blocks(shim):
    """Ragged pick lists are re-laid out per block of events over the union of their tables (host_logic.hpp):
    every event keeps its used picks, in its own order; padding picks are unused (-1); a block that is already
    uniform, or that holds an event not in increasing table order, is copied unchanged."""
    rng = np.random.default_rng(1)
    eb, ntab = 4, 12
    optr, tid, tobs, var = [0], [], [], []
    lists = []
    for e in range(11):
        k = 0 if e == 2 else int(rng.integers(1, ntab + 1))
        ids = np.sort(rng.permutation(ntab)[:k])
        if e in (4, 5, 6, 7):                    # block 1 is uniform already (same tables in every slot, one masked pick)
            ids = np.array([1, 3, 8])
        if e == 9:                               # block 2 holds an event in decreasing order -> untouched
            ids = ids[::-1]
        used = [int(t) if rng.random() > 0.2 else -1 for t in ids]
        lists.append([(u, float(100 * e + j), 0.25 + j) for j, u in enumerate(used)])
        for u, a, b in lists[-1]:
            tid.append(u); tobs.append(a); var.append(b)
        optr.append(len(tid))
    o2, t2, a2, b2 = _align(shim, eb, optr, tid, tobs, var)
    for e in range(11):
        got = [(int(t2[p]), a2[p], b2[p]) for p in range(o2[e], o2[e + 1])]
        want_used = [x for x in lists[e] if x[0] >= 0]
        assert [x for x in got if x[0] >= 0] == want_used            # used picks: same values, same order
    for b0 in (0, 4, 8):
        blk = range(b0, min(b0 + eb, 11))
        if b0 == 0:                              # aligned: same length, same table (or unused) in every slot
            lens = {o2[e + 1] - o2[e] for e in blk}
            assert len(lens) == 1
            for j in range(lens.pop()):
                ids = {int(t2[o2[e] + j]) for e in blk} - {-1}
                assert len(ids) <= 1
            slots = [max(int(t2[o2[e] + j]) for e in blk) for j in range(o2[1] - o2[0])]
            assert slots == sorted(slots) and -1 not in slots   # the union, in increasing table order
        else:                                    # copied unchanged
            for e in blk:
                got = [(int(t2[p]), a2[p], b2[p]) for p in range(o2[e], o2[e + 1])]
                assert got == lists[e]


@pytest.mark.parametrize("nf,nf0,stagger_on", [(5, 3, True), (4, 4, False), (1, 1, False), (13, 7, True)])
def test_ticket_queue_is_a_valid_order(shim, nf, nf0, stagger_on):
    """The ticket order of the streaming brick kernel (two field groups, the second `stagger` brick levels behind):
    every (sweep, brick, field) task appears exactly once, and along the queue each field's tasks never go back
    in (sweep, level) -- the only order its dependencies need (same field, earlier sweep or lower level)."""
    nbx, nby, nbz = 3, 4, 2
    nl = nbx + nby + nbz - 2
    ptr, order = [0], []
    for d in range(nl):                      # BrickPlan::build: bricks sorted by I + J + K
        for K in range(nbz):
            for J in range(nby):
                I = d - K - J
                if 0 <= I < nbx:
                    order.append((I, J, K))
        ptr.append(len(order))
    nbricks = len(order)
    assert nbricks == nbx * nby * nbz
    stagger = nl // 2 if stagger_on else 0
    bl = np.array(ptr, np.int32)
    vptr = np.zeros(8 * nl + stagger + 1, np.int64)
    p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    shim.shim_ticket_table.restype = C.c_longlong
    total = shim.shim_ticket_table(nl, p(bl, C.c_int), nf, nf0, stagger, p(vptr, C.c_longlong))
    assert total == 8 * nbricks * nf and np.all(np.diff(vptr) >= 0)
    seen = set()
    last = {}
    out = np.zeros(4, np.int32)
    for t in range(total):
        shim.shim_decode_ticket(C.c_longlong(t), p(vptr, C.c_longlong), p(bl, C.c_int), nl, stagger, nf0, nf - nf0, p(out, C.c_int))
        s, lev, bidx, f = (int(x) for x in out)
        assert 0 <= s < 8 and 0 <= lev < nl and 0 <= bidx < ptr[lev + 1] - ptr[lev] and 0 <= f < nf
        key = (s, ptr[lev] + bidx, f)
        assert key not in seen
        seen.add(key)
        pos = s * nl + lev
        assert pos >= last.get(f, -1)
        last[f] = pos
    assert len(seen) == total


@pytest.fixture(scope="module")
def brick_emu(tmp_path_factory):
    O.build()
    exe = tmp_path_factory.mktemp("brick_emu") / "brick_emu"
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-o", str(exe),
                           os.path.join(ROOT, "tests", "brick_pipeline_emulation.c"), "-L", O.ORACLE_DIR, "-loracle", "-lm",
                           f"-Wl,-rpath,{O.ORACLE_DIR}"])
    return str(exe)


@pytest.mark.parametrize("args", [(24, 20, 22, 16, 1), (24, 20, 22, 16, 2), (16, 9, 40, 256, 3), (32, 24, 18, 8, 4),
                                  (8, 8, 30, 16, 5), (40, 17, 33, 32, 6)])
def test_brick_pipeline_schedule_reproduces_the_reference_order(brick_emu, args):
    """tests/brick_pipeline_emulation.c: the streaming brick schedule of fsm_bricks16.cu (ring slots loaded 6 steps
    ahead, write-back 4 steps behind, upwind y neighbour By + 5 and x neighbour 7 steps ahead, random interleaving of
    the bricks) gives the reference-ordered oracle bit for bit and stays inside the 11-slot ring."""
    out = subprocess.run([brick_emu] + [str(a) for a in args], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("MATCH"), out.stdout


@pytest.mark.parametrize("args", [(24, 20, 22, 16, 1), (16, 9, 40, 256, 3), (32, 24, 18, 8, 4), (8, 8, 30, 16, 5),
                                  (40, 17, 33, 32, 6)])
def test_brick_pipeline_schedule_of_the_blocked_layout(brick_emu, args):
    """The kernel's default data path: the x halo comes out of face copies that the owning brick stores in half planes
    11 / 12 and 15 / 16 steps after a plane was entered (whole planes in brick columns cut by the grid's y face), and
    the upwind x neighbour is By + 3 (cut columns: By + 7) steps ahead.  Every face value stored is final, every
    upwind face value read is this sweep's, own / downwind ones are still the previous sweep's, and the fields equal
    the oracle bit for bit for random interleavings."""
    out = subprocess.run([brick_emu] + [str(a) for a in args] + ["0", "0", "0", "1"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("MATCH"), out.stdout


def test_face_copy_lead_of_the_blocked_layout_is_tight(brick_emu):
    """One step less lead and the downwind x neighbour reads face values that have not been stored yet."""
    out = subprocess.run([brick_emu, "24", "16", "22", "16", "1", "-1", "0", "0", "1"], capture_output=True, text=True)
    assert out.returncode == 1 and out.stdout.startswith("MISMATCH"), out.stdout
    assert int(out.stdout.split("stale=")[1].split()[0]) > 0 and "not_final=0" in out.stdout


def test_brick_pipeline_leads_are_tight(brick_emu):
    """One step less lead on the upwind neighbours and the same model reads halo values too early."""
    out = subprocess.run([brick_emu, "24", "20", "22", "16", "1", "-1"], capture_output=True, text=True)
    assert out.returncode == 1 and out.stdout.startswith("MISMATCH"), out.stdout


def test_brick_skipping_rule_is_sound_in_the_pipelined_model(brick_emu):
    """Study for a later round: skipping a brick sweep whose inputs cannot have changed keeps
    the oracle's bits in the pipelined, overlapped-sweep model."""
    out = subprocess.run([brick_emu, "24", "20", "22", "16", "1", "0", "1"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("MATCH") and "skipped=" in out.stdout, out.stdout
    assert int(out.stdout.split("skipped=")[1].split()[0]) > 0
