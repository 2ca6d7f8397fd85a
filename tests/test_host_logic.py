"""Host-side scheduling logic that needs no GPU: launch ranges, bench workload split."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_ranges_cover_the_slab_once():
    from pystella_b200.fused import _launch_ranges
    for Nx in (4, 16, 31, 32, 33, 64, 200):
        for h in (1, 2, 4):
            for chunk in (16, 64):
                assert _launch_ranges(Nx, h, False, chunk) == [(0, Nx)]
                ranges = _launch_ranges(Nx, h, True, chunk)
                flat = []
                for r in ranges:
                    flat += list(r) if isinstance(r[0], tuple) else [r]
                covered = np.zeros(Nx, int)
                for b, e in flat:
                    assert 0 <= b < e <= Nx
                    covered[b:e] += 1
                assert (covered == 1).all(), (Nx, h, chunk, ranges)
                # the faces (outermost h planes on either side) are in the FIRST launch
                first = ranges[0]
                first = list(first) if isinstance(first[0], tuple) else [first]
                in_first = np.zeros(Nx, bool)
                for b, e in first:
                    in_first[b:e] = True
                assert in_first[:min(h, Nx)].all() and in_first[-min(h, Nx):].all()


def test_bench_workload_partition():
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    c3 = dict(bench.CONFIGS["c3"], name="c3")
    for world in (1, 2, 3, 8):
        ws = [bench.workload(c3, world, r) for r in range(world)]
        assert sum(w["rank_shape"][0] for w in ws) == 512
        assert all(w["rank_shape"][1:] == (512, 512) for w in ws)
        assert [w["x_start"] for w in ws] == list(np.cumsum([0] + [w["rank_shape"][0]
                                                                  for w in ws[:-1]]))
    c5 = dict(bench.CONFIGS["c5"], name="c5")
    w = bench.workload(c5, 8, 3)
    assert w["grid_shape"] == (1024, 1024, 1024) and w["rank_shape"] == (128, 1024, 1024)
    assert w["np_dtype"] == np.float32 and w["dx"] == (5 / 1024,) * 3
    c4 = dict(bench.CONFIGS["c4"], name="c4")
    w4 = bench.workload(c4, 8, 0)
    assert w4["gw"] and w4["nscalars"] == 2 and w4["rank_shape"] == (64, 512, 512)
    assert bench.config_dict(c4, w4, 8) == bench.config_dict(c4, bench.workload(c4, 8, 5), 8)
    small = dict(c3, grid=(16, 16, 16))
    f, d = bench.initial_fields(bench.workload(small, 1, 0))
    assert f.shape == (1, 20, 20, 20) and abs(f.mean() - 0.193) < 1e-5
    # the potential callable and its monomial table describe the same function
    V = bench.potential_callable(w4)
    x = [0.3, -0.7]
    ref = sum(c * x[0]**p[0] * x[1]**p[1] for c, p in bench.potential_terms(w4))
    assert abs(V(x) - ref) < 1e-12 * abs(ref)


def test_reference_arm_prints_contract_line():
    """bench.py --impl reference on a tiny lattice: one JSON line with the agreed keys."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl",
                          "reference", "--grid", "32", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline",
                "e2e"):
        assert key in line
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port"
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["steps"] == 2 and line["config"]["grid_shape"] == [32, 32, 32]
    # both arms describe the workload with the same keys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    cfg = dict(bench.CONFIGS["c3"], name="c3", grid=(32, 32, 32))
    assert line["config"] == bench.config_dict(cfg, bench.workload(cfg, 1, 0), 1)


def test_bench_extras_are_bounded_and_never_fatal(monkeypatch):
    """bench.run_extras: every extra record runs in its own subprocess; a timeout, a crash
    or garbage output costs that entry only, and the time budget is respected."""
    import json
    import subprocess
    import types
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    calls = []

    def fake_run(cmd, **kw):
        calls.append((cmd, kw))
        assert "--no-extras" in cmd and "--no-cpu" in cmd and kw["timeout"] <= 120
        if "--diagnostics" in cmd:
            return types.SimpleNamespace(returncode=0, stderr="", stdout=json.dumps(
                {"bin_power": {"ok": True}}) + "\n")
        name = cmd[cmd.index("--config") + 1]
        if name == "c2":
            line = {"value": 9.0e4, "unit": bench.UNIT, "steps": 40, "ms_per_step": 0.9,
                    "dtype": "f64", "config": {"workload": "w"},
                    "roofline": {"frac": 0.85, "kernel": "k"},
                    "parity": {"ok": True, "max_rel": 5e-14, "tol": 1e-12}}
            return types.SimpleNamespace(returncode=0, stderr="",
                                         stdout="noise\n" + json.dumps(line) + "\n")
        if name == "c5":
            raise subprocess.TimeoutExpired(cmd, kw["timeout"])
        return types.SimpleNamespace(returncode=-11, stderr="segfault", stdout="not json")

    monkeypatch.setattr(subprocess, "run", fake_run)
    args = types.SimpleNamespace(extras_budget=240.0)
    out = bench.run_extras(args)
    assert out["c2"]["value"] == 9.0e4 and out["c2"]["parity_ok"] is True
    assert out["c2"]["roofline_frac"] == 0.85
    assert out["c5"] == {"error": "timeout"}
    assert out["c4"]["error"] == "rc=-11"
    assert out["diagnostics"]["bin_power"]["ok"] is True
    assert len(calls) == 4
    # no budget left: nothing is launched
    calls.clear()
    out = bench.run_extras(types.SimpleNamespace(extras_budget=5.0))
    assert not calls and all("skipped" in out[k] for k in ("c2", "c5", "c4", "diagnostics"))


# ---------------------------------------------------------------------------------------
# index arithmetic of the warp-specialised pipeline (csrc/fused_stage_ws.cuh, WsCfg and the
# consumer's own_off / pw_off / k), restated and checked by enumeration: every interior
# site is owned exactly once, every stencil read lands on the shared-memory column that
# holds the wanted padded coordinate, TMA boxes start 16-byte aligned, vector accesses are
# aligned in shared and global memory, and the periodic z images go to the right halo cells.
# ---------------------------------------------------------------------------------------
import math  # noqa: E402

def _ws_cfg(es, H, VEC, TZ, AL):
    SH = H % VEC if AL else 0
    GV = VEC if AL else math.gcd(VEC, H)
    BZ = TZ * VEC
    HL = -(-H // GV) * GV
    WQ = 16 // es
    W = -(-(BZ + 2 * HL) // WQ) * WQ
    PWL = 0 if AL else H % WQ
    PWW = -(-(BZ + PWL) // WQ) * WQ
    ZF = H - HL - SH
    ZP = H - PWL - SH
    return dict(SH=SH, GV=GV, BZ=BZ, HL=HL, W=W, PWL=PWL, PWW=PWW, ZF=ZF, ZP=ZP, WQ=WQ)

def _ws_check(es, H, VEC, TZ, Nz, AL):
    c = _ws_cfg(es, H, VEC, TZ, AL)
    Pz = Nz + 2 * H
    ntiles = -(-(Nz + c["SH"]) // c["BZ"])
    written = {}
    halo_written = {}
    for b in range(ntiles):
        z0 = b * c["BZ"]
        # TMA start coordinates must be 16-byte aligned
        assert ((z0 + c["ZF"]) * es) % 16 == 0, ("f box", z0, c)
        assert ((z0 + c["ZP"]) * es) % 16 == 0, ("pw box", z0, c)
        for tz in range(TZ):
            k = z0 + tz * VEC - c["SH"]
            own = c["HL"] + tz * VEC
            pw = c["PWL"] + tz * VEC
            # vector accesses: alignment in shared memory and in global memory
            assert (own * es) % (c["GV"] * es) == 0 and (pw * es) % (c["GV"] * es) == 0
            assert ((k + H) % c["GV"]) == 0, ("global alignment", k, H, c)
            assert own - c["HL"] >= 0 and own + VEC + c["HL"] <= c["W"]
            assert pw + VEC <= c["PWW"]
            for v in range(VEC):
                z = k + v
                ok = 0 <= z < Nz
                if not AL:
                    ok = k < Nz
                    if ok:
                        assert z < Nz, "non-AL needs Nz % VEC == 0"
                if not ok:
                    continue
                # smem column <-> padded coordinate
                assert z0 + c["ZF"] + own + v == z + H
                assert z0 + c["ZP"] + pw + v == z + H
                for s in range(-H, H + 1):
                    col = own + v + s
                    assert 0 <= col < c["W"]
                    pz = z0 + c["ZF"] + col
                    assert 0 <= pz < Pz, (pz, Pz)          # never an out-of-bounds column
                assert z not in written
                written[z] = (b, tz, v)
                # fout pointer is at padded (k + H); element v; wrap images
                base = k + H
                if z < H:
                    halo_written[base + v + Nz] = z
                if z >= Nz - H:
                    halo_written[base + v - Nz] = z
    assert sorted(written) == list(range(Nz)), "every interior site exactly once"
    for pz, z in halo_written.items():
        assert (pz < H or pz >= Nz + H) and (pz - H) % Nz == z
    assert len(halo_written) == 2 * H
    return ntiles



def test_pipeline_index_arithmetic_aligned_mapping():
    """The experimental aligned mapping (template flag AL, tune_variant 16)."""
    for es, VEC in ((4, 4), (8, 2)):
        for H in (1, 2, 3, 4):
            for TZ in (16, 32):
                for Nz in range(max(4, H), 200):
                    if ((Nz + 2 * H) * es) % 16:
                        continue
                    _ws_check(es, H, VEC, TZ, Nz, True)


def test_pipeline_index_arithmetic_default_mapping():
    """The mapping of the validated kernels, for the layouts it is dispatched for (consecutive
    sites per thread: GV == VEC, i.e. fp64 and fp32 with radius 4)."""
    for es, VEC, TZ in ((8, 2, 64), (4, 2, 64), (4, 4, 32)):
        for H in (2, 4):
            if math.gcd(VEC, H) != VEC:
                continue          # split mapping: test below
            for Nz in range(8, 200, 2):
                if ((Nz + 2 * H) * es) % 16 or Nz % VEC:
                    continue
                _ws_check(es, H, VEC, TZ, Nz, False)


def _ws_check_split(es, H, VEC, TZ, Nz):
    """Split vectors (csrc/fused_stage_ws.cuh, SPLIT: GV = gcd(VEC, H) < VEC): a thread's VEC
    sites are VEC/GV pieces of GV consecutive sites, TZ*GV apart.  Every interior site is owned
    exactly once, pieces are GV-aligned in shared and global memory, the lanes of a warp touch
    consecutive pieces (one contiguous span per vector access), stencil reads stay inside the
    tile, and the periodic z images land on the right halo cells."""
    c = _ws_cfg(es, H, VEC, TZ, False)
    GV, BZ, HL = c["GV"], c["BZ"], c["HL"]
    assert GV < VEC
    NG, GS = VEC // GV, TZ * GV
    Pz = Nz + 2 * H
    ntiles = -(-Nz // BZ)
    written, halo_written = {}, {}
    for b in range(ntiles):
        z0 = b * BZ
        assert ((z0 + c["ZF"]) * es) % 16 == 0 and ((z0 + c["ZP"]) * es) % 16 == 0
        for g in range(NG):
            starts = []
            for tz in range(TZ):
                k = z0 + tz * GV                       # thread's first site
                own = HL + tz * GV + g * GS            # smem column of piece g
                pw = c["PWL"] + tz * GV + g * GS
                starts.append(own)
                assert own % GV == 0 and pw % GV == 0 and (k + g * GS + H) % GV == 0
                assert own - HL >= 0 and own + GV + HL <= c["W"] and pw + GV <= c["PWW"]
                if not (k + g * GS < Nz):              # pact[g]
                    continue
                for e in range(GV):
                    z = k + g * GS + e
                    assert z < Nz, "needs Nz % GV == 0"
                    assert z0 + c["ZF"] + own + e == z + H
                    assert z0 + c["ZP"] + pw + e == z + H
                    for s_ in range(-H, H + 1):
                        col = own + e + s_
                        assert 0 <= col < c["W"] and 0 <= z0 + c["ZF"] + col < Pz
                    assert z not in written
                    written[z] = (b, tz, g, e)
                    voff = g * GS + e                  # element offset from the first site
                    base = k + H
                    if z < H:
                        halo_written[base + voff + Nz] = z
                    if z >= Nz - H:
                        halo_written[base + voff - Nz] = z
            # lanes of a warp (consecutive tz) read consecutive pieces: contiguous span
            assert all(b_ - a_ == GV for a_, b_ in zip(starts, starts[1:]))
    assert sorted(written) == list(range(Nz))
    for pz, z in halo_written.items():
        assert (pz < H or pz >= Nz + H) and (pz - H) % Nz == z
    assert len(halo_written) == 2 * H


def test_pipeline_index_arithmetic_split_mapping():
    """fp32, radius 2: 16-byte vectors as two 8-byte pieces 64 sites apart."""
    for TZ in (16, 32):
        for Nz in range(8, 300, 2):
            if ((Nz + 4) * 4) % 16:
                continue
            _ws_check_split(4, 2, 4, TZ, Nz)


# ---------------------------------------------------------------------------------------
# ring protocol of the warp-specialised pipeline (csrc/fused_stage_ws.cuh): a discrete
# simulation of the producer lane, the consumer warps and the asynchronous TMA completions
# under random scheduling.  Checks: no dead-lock, no phase aliasing (a wait never sees a
# barrier two phases ahead), every consumer reads the plane the producer loaded for it, and
# no slot is overwritten while a warp still needs it.
# ---------------------------------------------------------------------------------------

class _Mbar:
    def __init__(self, count):
        self.count, self.pending, self.completed = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.completed += 1
            self.pending = self.count

    def done(self, phase_index):
        """wait(parity of phase_index): passes once that phase has completed; the protocol
        must never let the barrier run two phases ahead of a waiter."""
        assert self.completed <= phase_index + 1, "phase aliasing"
        return self.completed > phase_index


def _simulate_rings(H, P, ncomp, ncw, rng):
    RF, RP = H + 1 + P, P + 1
    nsteps = ncomp + 2 * H
    full_f = [_Mbar(1) for _ in range(RF)]
    empty_f = [_Mbar(ncw) for _ in range(RF)]
    full_p = [_Mbar(1) for _ in range(RP)]
    empty_p = [_Mbar(ncw) for _ in range(RP)]
    fslot, pslot = [None] * RF, [None] * RP          # what the slots currently hold
    in_flight = []                                   # issued TMA loads: (kind, slot, tag)
    prod = {"m": 0, "stage": 0}
    cons = [{"n": 0, "stage": 0} for _ in range(ncw)]

    def producer_step():
        m = prod["m"]
        if m >= nsteps:
            return False
        if prod["stage"] == 0:
            slot = m % RF
            if m >= RF and not empty_f[slot].done(m // RF - 1):
                return False
            in_flight.append(("f", slot, m))
            prod["stage"] = 1
            return True
        c = m - 2 * H
        if 0 <= c < ncomp:
            slot = c % RP
            if c >= RP and not empty_p[slot].done(c // RP - 1):
                return False
            in_flight.append(("p", slot, c))
        prod["m"], prod["stage"] = m + 1, 0
        return True

    def tma_step():
        if not in_flight:
            return False
        kind, slot, tag = in_flight.pop(rng.integers(len(in_flight)))
        if kind == "f":
            fslot[slot] = tag
            full_f[slot].arrive()
        else:
            pslot[slot] = tag
            full_p[slot].arrive()
        return True

    def consumer_step(w):
        st = cons[w]
        n = st["n"]
        if n >= nsteps:
            return False
        if st["stage"] == 0:                      # wait for f plane n, shift the x queue
            if not full_f[n % RF].done(n // RF):
                return False
            assert fslot[n % RF] == n
            st["stage"] = 1
            return True
        if st["stage"] == 1:                      # compute site plane c = n - 2H
            if n >= 2 * H:
                c = n - 2 * H
                if not full_p[c % RP].done(c // RP):
                    return False
                assert pslot[c % RP] == c
                assert fslot[(n - H) % RF] == n - H, "centre plane overwritten"
                empty_p[c % RP].arrive()
            st["stage"] = 2
            return True
        if n >= H:
            empty_f[(n - H) % RF].arrive()
        st["n"], st["stage"] = n + 1, 0
        return True

    idle = 0
    while prod["m"] < nsteps or any(c["n"] < nsteps for c in cons) or in_flight:
        actors = [producer_step, tma_step] + [lambda w=w: consumer_step(w) for w in range(ncw)]
        moved = actors[rng.integers(len(actors))]()
        idle = 0 if moved else idle + 1
        if idle > 200 * (ncw + 2):
            # nobody moved for a long time: make sure it is not just bad luck
            assert any(a() for a in actors), "dead-lock"
            idle = 0


def test_pipeline_ring_protocol_is_deadlock_free():
    rng = np.random.default_rng(5)
    for H in (1, 2, 3, 4):
        for P in (2, 3):
            for ncomp in (1, 2, 5, 16):
                for ncw in (1, 4, 8):
                    _simulate_rings(H, P, ncomp, ncw, rng)


# ---------------------------------------------------------------------------------------
# persistent variant (csrc/fused_stage_pws.cuh): the same rings run CONTINUOUSLY across
# segments; slots carry headers; at the end of a segment the consumer releases the H trailing
# planes as well.  The simulation restates producer and consumer loops and checks, under
# random scheduling: no dead-lock, no phase aliasing, every consumer sees the header and the
# plane the producer meant for that step, centre planes are still intact when they are read,
# every slot use is released exactly once, every site plane of every segment is produced
# exactly once, and the guided schedule partitions the items identically for every run.
# ---------------------------------------------------------------------------------------

def _pws_schedule(Nx, ncol, lmax=128, phases=6):
    """csrc/fused_stage_pws.cuh::pws_make_schedule + pws_chunk, restated: chunk number ->
    (column, out_lo, out_hi)."""
    x, chunk, table = 0, 0, []
    for j in range(phases):
        ln = lmax >> j
        rem = Nx - x
        n = -(-rem // ln) if j == phases - 1 else ((rem - ln) // ln if rem > ln else 0)
        table.append((chunk, x, ln))
        x += rem if j == phases - 1 else n * ln
        chunk += n * ncol
    ends = [t[1] for t in table[1:]] + [Nx]
    out = []
    for i in range(chunk):
        ph = max(j for j, (c0, _, _) in enumerate(table) if i >= c0)
        c0, x0, ln = table[ph]
        local = i - c0
        lo = x0 + (local // ncol) * ln
        out.append((local % ncol, lo, min(lo + ln, ends[ph])))
    return out


def test_pws_schedule_is_a_static_partition():
    for Nx, ncol in ((64, 512), (512, 512), (128, 2048), (37, 12), (4, 3), (1, 1), (300, 7)):
        chunks = _pws_schedule(Nx, ncol)
        cover = np.zeros((ncol, Nx), int)
        for col, lo, hi in chunks:
            assert 0 <= lo < hi <= Nx
            cover[col, lo:hi] += 1
        assert (cover == 1).all()                     # every site plane exactly once
        lens = [hi - lo for _, lo, hi in chunks]
        assert max(lens) <= 128 and lens[-1] <= 4      # short chunks last
        # neighbouring chunk numbers = neighbouring columns of the same x range
        for (c0, lo0, hi0), (c1, lo1, hi1) in zip(chunks, chunks[1:]):
            if c1 == c0 + 1:
                assert (lo0, hi0) == (lo1, hi1)
        if Nx >= 256:
            assert lens[0] == 128


def _simulate_pws(H, P, segments, ncw, rng):
    """segments: list of (out_lo, out_hi) handled by one CTA, in order."""
    RF, RP = H + 1 + P, P + 1
    full_f = [_Mbar(1) for _ in range(RF)]
    empty_f = [_Mbar(ncw) for _ in range(RF)]
    full_p = [_Mbar(1) for _ in range(RP)]
    empty_p = [_Mbar(ncw) for _ in range(RP)]
    fslot, pslot, hdr = [None] * RF, [None] * RP, [None] * RF
    in_flight = []
    # the producer's plan: one entry per f plane (+ the terminate marker)
    plan = []
    for si, (lo, hi) in enumerate(segments):
        nseg = (hi - lo) + 2 * H
        for m in range(nseg):
            plan.append(dict(seg=si, xpad=lo + m, lo=lo, hi=hi, end=(m == nseg - 1), term=False))
    plan.append(dict(term=True))
    prod = {"g": 0, "gp": 0, "stage": 0}
    cons = [{"g": 0, "gp": 0, "done": False} for _ in range(ncw)]
    produced = []

    def producer_step():
        g = prod["g"]
        if g >= len(plan):
            return False
        e = plan[g]
        slot = g % RF
        if prod["stage"] == 0:
            if g >= RF and not empty_f[slot].done(g // RF - 1):
                return False
            hdr[slot] = e
            if e["term"]:
                full_f[slot].arrive()
                prod["g"] = g + 1
                return True
            in_flight.append(("f", slot, g))
            prod["stage"] = 1
            return True
        m = e["xpad"] - e["lo"]
        if m >= 2 * H:
            gp = prod["gp"]
            ps_ = gp % RP
            if gp >= RP and not empty_p[ps_].done(gp // RP - 1):
                return False
            in_flight.append(("p", ps_, (e["seg"], e["xpad"] - 2 * H)))
            prod["gp"] = gp + 1
        prod["g"], prod["stage"] = g + 1, 0
        return True

    def tma_step():
        if not in_flight:
            return False
        kind, slot, tag = in_flight.pop(rng.integers(len(in_flight)))
        if kind == "f":
            fslot[slot] = tag
            full_f[slot].arrive()
        else:
            pslot[slot] = tag
            full_p[slot].arrive()
        return True

    def consumer_step(w):
        c = cons[w]
        if c["done"]:
            return False
        g = c["g"]
        slot = g % RF
        if not full_f[slot].done(g // RF):
            return False
        e = hdr[slot]
        assert e is plan[g], "header of another step"
        if e["term"]:
            c["done"] = True
            return True
        assert fslot[slot] == g, "own plane overwritten"
        m = e["xpad"] - e["lo"]
        if m >= 2 * H:
            gp = c["gp"]
            ps_ = gp % RP
            if not full_p[ps_].done(gp // RP):
                return False
            assert fslot[(g - H) % RF] == g - H, "centre plane overwritten"
            assert pslot[ps_] == (e["seg"], e["xpad"] - 2 * H), "wrong pointwise plane"
            if w == 0:
                produced.append((e["seg"], e["xpad"] - 2 * H))
            empty_p[ps_].arrive()
            c["gp"] = gp + 1
        if m >= H:
            empty_f[(g - H) % RF].arrive()
        if e["end"]:
            for s in range(H - 1, -1, -1):
                empty_f[(g - s) % RF].arrive()
        c["g"] = g + 1
        return True

    idle = 0
    while not all(c["done"] for c in cons):
        actors = [producer_step, tma_step] + [lambda w=w: consumer_step(w) for w in range(ncw)]
        if actors[rng.integers(len(actors))]():
            idle = 0
        else:
            idle += 1
            assert idle < 20000, "dead-lock"
    want = [(si, i) for si, (lo, hi) in enumerate(segments) for i in range(lo, hi)]
    assert produced == want
    # every slot use released exactly once: the barriers are back in their initial state
    nplanes = len(plan) - 1
    for s in range(RF):
        uses = len(range(s, nplanes, RF))
        assert empty_f[s].completed == uses and empty_f[s].pending == ncw


def test_pws_ring_protocol_simulation():
    rng = np.random.default_rng(2024)
    for H, P in ((2, 3), (2, 2), (2, 1), (1, 2), (4, 1)):
        for _ in range(6):
            nseg = int(rng.integers(1, 5))
            segments = []
            for _s in range(nseg):
                lo = int(rng.integers(0, 40)) * 4
                segments.append((lo, lo + int(rng.integers(1, 9)) * 4 - int(rng.integers(0, 3))))
            _simulate_pws(H, P, segments, int(rng.integers(1, 5)), rng)
