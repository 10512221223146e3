"""Diagnostic runner for the tcgen05 convolution kernel: each layer case runs in its own process so a
device trap (pipeline watchdog) in one case cannot poison the others.  Prints error statistics and
the watchdog words.  Usage on the GPU box:  python tools/tc_probe.py  [--case N]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_case(i):
    import numpy as np
    import torch
    from test_gpu_fp16 import CASES, ref_layer, vp, stream
    from yolo_v3_b200 import _lib, synth, topology
    li, B, H, W, with_res = CASES[i]
    spec = topology.layer_specs(80)[li]
    sd = synth.make_state_dict(seed=1234, recipe="analytic")
    lib = _lib.load()
    ctx = _lib.create_ctx(0, 80, None)
    for k in ([spec["key"] + s for s in (".conv.weight", ".bn.weight", ".bn.bias", ".bn.running_mean", ".bn.running_var")]
              if spec["bn"] else [spec["key"] + ".weight", spec["key"] + ".bias"]):
        v = sd[k].contiguous()
        _lib.check(lib.yb_set_tensor(ctx, k.encode(), vp(v), v.numel(), 1), ctx)
    _lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP16), ctx)
    rs = np.random.RandomState(100 + li)
    x = torch.from_numpy(rs.standard_normal((B, H, W, spec["cin"])).astype(np.float32)).half()
    Ho, Wo = H // spec["stride"], W // spec["stride"]
    res = torch.from_numpy(rs.standard_normal((B, Ho, Wo, spec["cout"])).astype(np.float32)).half() if with_res else None
    head = not spec["bn"]
    cs = (spec["cout"] + 15) // 16 * 16 if head else spec["cout"]
    out = torch.full((B, Ho, Wo, cs), float("nan"), device="cuda", dtype=torch.float32 if head else torch.float16)
    xd = x.cuda()
    rd = res.cuda() if with_res else None
    tag = f"case {i}: layer {li} {spec['key']} cin={spec['cin']} cout={spec['cout']} ks={spec['ks']} s={spec['stride']} B={B} H={H} W={W} res={with_res}"
    try:
        _lib.check(lib.yb_run_layer(ctx, li, vp(xd), B, H, W, vp(rd) if with_res else None, vp(out), stream()), ctx)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        words = (ctypes.c_int * 8)()
        lib.yb_debug_words(ctx, words, 8)
        print(f"FAIL {tag}\n     exception: {str(e)[:300]}\n     watchdog words: {list(words)}")
        return 1
    y = out.float().cpu()[..., :spec["cout"]]
    ref = ref_layer(sd, spec, x, res)
    err = (y - ref).abs()
    tol = 3e-3 * ref.abs().max() + 2e-3 * ref.abs()
    bad = (err > tol) | torch.isnan(y)
    nb = int(bad.sum())
    print(f"{'ok  ' if nb == 0 else 'BAD '} {tag}: max err {float(err[~torch.isnan(err)].max()) if (~torch.isnan(err)).any() else float('nan'):.4g} "
          f"ref max {float(ref.abs().max()):.4g} bad {nb}/{bad.numel()} nan {int(torch.isnan(y).sum())}")
    if nb:
        M = B * Ho * Wo
        badm = bad.view(M, -1)
        rows_bad = badm.any(1)
        tiles = [int(rows_bad[t * 128:(t + 1) * 128].sum()) for t in range((M + 127) // 128)]
        print("     bad rows per 128-row tile:", tiles[:24])
        colchunks = [int(badm[:, c:c + 16].any(1).sum()) for c in range(0, badm.shape[1], 16)]
        print("     bad rows per 16-col chunk:", colchunks[:24])
        idx = bad.nonzero()[:6]
        for t in idx:
            t = tuple(int(v) for v in t)
            print(f"     at {t}: got {float(y[t]):.5f} want {float(ref[t]):.5f}")
        # does the output match the reference of a shifted pixel? (im2col coordinate bugs)
        yv, rv = y.view(M, -1), ref.view(M, -1)
        for sh in (1, -1, Wo, -Wo):
            a, b = (yv[sh:], rv[:-sh]) if sh > 0 else (yv[:sh], rv[-sh:])
            print(f"     mean|y[m+{sh}]-ref[m]| = {float((a - b).abs().mean()):.4f}", end=";")
        print(f" mean|y-ref| = {float((yv - rv).abs().nan_to_num(9).mean()):.4f}, mean|ref| = {float(rv.abs().mean()):.4f}")
    return 1 if nb else 0


if __name__ == "__main__":
    if "--case" in sys.argv:
        sys.exit(run_case(int(sys.argv[sys.argv.index("--case") + 1])))
    from test_gpu_fp16 import CASES
    fails = 0
    for i in range(len(CASES)):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i)], capture_output=True, text=True, timeout=180)
            sys.stdout.write(r.stdout)
            if r.returncode != 0 and not r.stdout.strip():
                sys.stdout.write(f"FAIL case {i}: rc={r.returncode} stderr tail: {r.stderr[-400:]}\n")
            fails += r.returncode != 0
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT case {i}")
            fails += 1
    print(f"tc_probe: {len(CASES) - fails}/{len(CASES)} cases ok")
