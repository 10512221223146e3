"""CPU-only tests of the host side: the C-ABI library loads and exports every declared symbol, the
Python mirror keeps the reference's names/signatures, fails loudly without a GPU, and the multi-GPU
host logic works over gloo with world_size 2."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from yolo_v3_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "yolo_v3_b200", "csrc"), "-j8"])
    return _lib


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "yolo_b200.h")).read()
    declared = set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/yolo_b200.h but not exported"
    assert declared == set(built_lib.PROTOTYPES), "ctypes prototypes and header disagree"
    built_lib.load()


def test_library_is_sm100a_tensor_core_code(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "IM2COL"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built_lib):
    from yolo_v3_b200 import YoloNet, postprocessing
    with pytest.raises(built_lib.YbError) as e:
        built_lib.create_ctx(0, 80, None)
    assert "no CPU fallback" in str(e.value)
    net = YoloNet((64, 64)).eval()
    with pytest.raises(RuntimeError):
        net(torch.rand(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        postprocessing(torch.rand(1, 10, 85), 80)


def test_mirror_keeps_reference_names_and_signatures(golden):
    from yolo_v3_b200 import YoloNet, postprocessing, synth
    g = golden("net_golden.npz")
    net = YoloNet((416, 416))
    keys = list(net.state_dict().keys())
    assert keys == [str(k) for k in g["state_dict_keys"]]          # the reference's own state_dict order
    shapes = [tuple(v.shape) for v in net.state_dict().values()]
    assert shapes == [tuple(int(x) for x in s.split(",") if x) for s in g["state_dict_shapes"]]
    assert list(inspect.signature(YoloNet.forward).parameters) == ["self", "x", "target"]
    assert list(inspect.signature(YoloNet.loadWeight).parameters) == ["self", "weights_path", "format"]
    assert list(inspect.signature(postprocessing).parameters)[:6] == ["detections", "num_classes", "obj_conf_thr",
                                                                      "nms_thr", "is_eval", "use_nms"]
    p = inspect.signature(postprocessing).parameters
    assert (p["obj_conf_thr"].default, p["nms_thr"].default, p["is_eval"].default, p["use_nms"].default) == (0.5, 0.4, False, True)
    assert net.numClass == 80 and net.img_dim == (416, 416) and len(net.stat_keys) == 10
    assert [n for n, _ in net.named_children()] == ["feature", "pre_det1", "yolo1", "up1", "pre_det2", "yolo2", "up2",
                                                    "pre_det3", "yolo3"]
    sd = synth.make_state_dict(recipe="analytic")
    net.load_state_dict(sd)
    assert sum(p.numel() for p in net.parameters()) == 61949149


def test_python_darknet_loader_roundtrip(tmp_path, oracle):
    from yolo_v3_b200 import YoloNet, synth
    sd = synth.make_state_dict(recipe="analytic")
    blob = oracle.darknet_blob_from_state_dict(sd)
    path = tmp_path / "synth.weights"
    with open(path, "wb") as fp:
        np.array([0, 2, 0, 1234, 0], np.int32).tofile(fp)
        blob.tofile(fp)
    net = YoloNet((64, 64))
    net.loadWeight(str(path), "darknet")
    assert int(net.seen) == 1234
    assert all(torch.equal(v, sd[k]) for k, v in net.state_dict().items() if "num_batches" not in k)
    # backbone-only stream through net.feature.loadWeight (darknet.py:102-104)
    bb = tmp_path / "bb.weights"
    with open(bb, "wb") as fp:
        np.array([0, 2, 0, 0, 0], np.int32).tofile(fp)
        oracle.darknet_blob_from_state_dict(sd, backbone_only=True).tofile(fp)
    net2 = YoloNet((64, 64))
    net2.feature.loadWeight(str(bb))
    s2 = net2.state_dict()
    assert all(torch.equal(s2[k], sd[k]) for k in sd if k.startswith("feature.") and "num_batches" not in k)
    with pytest.raises(ValueError):
        net2.load_darknet_stream(blob[:1000])
    out = tmp_path / "pt.pth"
    net.saveWeight(str(out))
    net3 = YoloNet((64, 64))
    net3.loadWeight(str(out))
    assert all(torch.equal(v, sd[k]) for k, v in net3.state_dict().items() if "num_batches" not in k)


def test_shard_bounds_and_merge():
    from yolo_v3_b200.parallel import merge_gathered, shard_bounds
    for gb, world in ((256, 8), (32, 4), (10, 4), (3, 8)):
        spans = [shard_bounds(gb, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)
    rows = torch.arange(4 * 3 * 7, dtype=torch.float32).view(4, 3, 7)
    out = merge_gathered(rows, torch.tensor([2, 0, 3, 1]))
    assert [tuple(o.shape) for o in out] == [(2, 7), (0,), (3, 7), (1, 7)]
    assert torch.equal(out[2], rows[2])
    assert merge_gathered(rows, torch.zeros(4, dtype=torch.int32), cand_any=False) == []


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolo_v3_b200.parallel import exchange_unique_id, merge_gathered, shard_bounds
    uid = exchange_unique_id(lambda: bytes(range(128)), rank)
    # emulate the per-batch gather of fixed-capacity rows with the host collective
    lo, hi = shard_bounds(6, world, rank)
    cap = 4
    rows = torch.zeros(hi - lo, cap, 7)
    counts = torch.zeros(hi - lo, dtype=torch.int32)
    for i, img in enumerate(range(lo, hi)):
        counts[i] = img % (cap + 1)
        rows[i, :counts[i]] = float(img)
    all_rows = [torch.zeros_like(rows) for _ in range(world)]
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_rows, rows)
    dist.all_gather(all_counts, counts)
    merged = merge_gathered(torch.cat(all_rows), torch.cat(all_counts))
    ok = uid == bytes(range(128)) and len(merged) == 6
    for img, m in enumerate(merged):
        ok = ok and len(m) == img % (cap + 1) and (m.numel() == 0 or bool((m == float(img)).all()))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_host_helpers_of_the_next_rows(oracle):
    """Pure-host pieces of the pre-/post-path mirrors: letterbox geometry and the evaluate.py result entries."""
    import json
    from yolo_v3_b200.evaluate import create_results_entry, get_image_id_from_path
    from yolo_v3_b200.utils import letterbox_transforms
    for inner, outer in (((602, 452), (416, 416)), ((333, 500), (416, 416)), ((1280, 720), (608, 608)), ((64, 64), (96, 96))):
        assert letterbox_transforms(inner, outer) == oracle.letterbox_transforms(inner, outer)
    assert letterbox_transforms((602, 452), (416, 416))[:4] == (416, 312, 0, 52)           # SURVEY.md 8c fixture
    assert get_image_id_from_path("coco/images/val2014/COCO_val2014_000000000139.jpg") == 139
    e = create_results_entry(139, 17, [1.5, 2.0, 30.0, 40.25], 0.875)
    assert json.dumps(e, separators=(",", ":")) == '{"image_id":139,"category_id":17,"bbox":[1.5,2.0,30.0,40.25],"score":0.875}'


def test_named_tensors_follow_the_live_state_dict():
    """YoloNet._named_tensors feeds the change signature checked before every forward; it reads the modules' own
    parameter / buffer tables through slots resolved once, and must always yield exactly the live state_dict tensors."""
    from yolo_v3_b200 import YoloNet

    def same(net):
        a = list(net._named_tensors())
        b = list(net.state_dict(keep_vars=True).items())
        return len(a) == len(b) == 438 and all(k1 == k2 and t1 is t2 for (k1, t1), (k2, t2) in zip(a, b))

    net = YoloNet((416, 416))
    dev = torch.device("cpu")
    assert same(net)
    s0 = net._signature(dev)
    assert net._signature(dev) == s0                      # nothing changed
    sd = {k: v.clone() + 1 for k, v in net.state_dict().items()}
    net.load_state_dict(sd)                               # in-place copies: same tensors, new versions
    s1 = net._signature(dev)
    assert same(net) and s1 != s0
    net.double()                                          # Module._apply replaces buffers and parameter data
    assert same(net) and net._signature(dev) != s1
    net.feature.float()                                   # ... also when applied to a sub-module only
    assert same(net)
    with torch.no_grad():
        net.pre_det1.mlist[6].bias.add_(1.0)              # a direct in-place edit of one tensor
    assert net._signature(dev) != s1 and same(net)


def test_data_writes_need_refresh_weights_or_check_weights():
    """`param.data.copy_()` (the idiom of the reference's WeightManager, darknet.py:275) does not bump the tensor's version
    counter: the default signature cannot see it (documented), refresh_weights() forgets the upload, and
    check_weights=True sees the new contents."""
    from yolo_v3_b200 import YoloNet
    dev = torch.device("cpu")
    net = YoloNet((416, 416))
    s0 = net._signature(dev)
    net._sig = s0                                          # as after a forward
    p = net.feature.mlist[0].conv.weight
    p.data.mul_(2.0)
    net.feature.mlist[0].bn.running_mean.data.copy_(torch.ones(32))
    assert net._signature(dev) == s0                      # the limitation
    net.refresh_weights()
    assert net._sig is None                                # next forward re-uploads
    chk = YoloNet((416, 416), check_weights=True)
    c0 = chk._signature(dev)
    assert chk._signature(dev) == c0
    chk.feature.mlist[0].conv.weight.data.mul_(2.0)
    c1 = chk._signature(dev)
    assert c1 != c0
    chk.pre_det3.mlist[6].bias.data.copy_(torch.full((255,), 0.25))
    assert chk._signature(dev) != c1
    # load_state_dict / _apply / the darknet loader invalidate explicitly, whatever the version counters say
    net._sig = s0
    net.load_state_dict(net.state_dict())
    assert net._sig is None
    net._sig = s0
    net.float()
    assert net._sig is None


def test_copy_and_pickle_drop_the_engine_and_rebind_the_backbone():
    """copy.deepcopy / pickle of a YoloNet: the copy has no engine context of its own yet, and its `feature` points back at
    the COPY (loadWeight / forward on copy.feature must not act on the original)."""
    import copy
    import io
    import pickle
    from yolo_v3_b200 import YoloNet
    net = YoloNet((416, 416), numClass=20)
    net._ctx = ctypes.c_void_p(12345)                      # stands for a live engine (never dereferenced here)
    net._sig = ("x",)
    try:
        cp = copy.deepcopy(net)
        assert cp._ctx is None and cp._sig is None and cp.feature._owner() is cp and net.feature._owner() is net
        assert cp.numClass == 20 and all(torch.equal(a, b) for a, b in zip(cp.state_dict().values(), net.state_dict().values()))
        rt = pickle.loads(pickle.dumps(net))
        assert rt._ctx is None and rt.feature._owner() is rt
        buf = io.BytesIO()
        torch.save(net, buf)
        buf.seek(0)
        back = torch.load(buf, weights_only=False)
        assert back._ctx is None and back.feature._owner() is back
    finally:
        net._ctx = None                                    # nothing to destroy


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference runs on the host cores only -- the reference's own modules from oracle/_ref where that
    directory exists (`make -C oracle ref`), else the oracle port: one JSON line on stdout with the keys the driver reads,
    whatever else the libraries print."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--size", "416"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "darknet.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("yolov3_416x416")


def test_reference_in_oracle_ref_is_what_the_oracle_restates():
    """Where oracle/_ref exists (build container: copied from /root/reference by `make -C oracle ref`; GPU box: shipped with
    the snapshot), the reference's own forward + postprocessing and the oracle port agree on a seeded input: logits to
    fp32 round-off of the same torch build, detections identical."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not present")
    import oracle.yolo_oracle as O
    from yolo_v3_b200 import synth
    darknet, utils = ref_loader.load()
    assert "utils" not in sys.modules or not getattr(sys.modules["utils"], "__file__", "").startswith(ref_loader.REF_DIR)
    sd = synth.make_state_dict(seed=1234, recipe="calibrated")
    net = darknet.YoloNet((160, 96))
    net.load_state_dict(sd)
    net.eval()
    x = synth.make_images(2, 96, 160, seed=4)
    with ref_loader.cpu_only(), torch.no_grad():
        det = torch.cat(net(x, None), 1)
        res = utils.postprocessing(det.clone(), 80, 0.3, 0.4)
    ref = torch.cat(O.forward(sd, x), 1)
    assert torch.allclose(det, ref, rtol=1e-5, atol=1e-6)
    mine = O.postprocessing(ref, 80, 0.3, 0.4)
    assert len(res) == len(mine) and all(torch.equal(a, b) for a, b in zip(res, mine))


def test_bench_detection_matching():
    """bench.py's parity block matches two detection lists one to one (same class, IOU > 0.5): identical lists match
    completely, a shifted box still matches with its IOU, a class change or a far box does not."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    ref = torch.tensor([[10., 10., 50., 50., .9, .8, 3.], [100., 100., 180., 160., .7, .6, 3.], [30., 200., 90., 260., .8, .7, 17.]])
    m, ng, nr, miou, ds = b.match_detections(ref.clone(), ref)
    assert (m, ng, nr) == (3, 3, 3) and abs(miou - 1.0) < 1e-6 and ds == 0.0
    got = ref.clone()
    got[0, :4] += torch.tensor([2., 0., 2., 0.])          # shifted by 2 px: IOU 38*40 / (2*1600 - 1520)
    got[1, 6] = 4.                                        # other class: no match
    got[2, 5] = 0.65
    m, ng, nr, miou, ds = b.match_detections(got, ref)
    assert (m, ng, nr) == (2, 3, 3) and abs(ds - 0.05) < 1e-6
    assert abs(miou - (1520. / 1680. + 1.0) / 2) < 1e-6
    assert b.match_detections(torch.zeros(0, 7), ref)[:3] == (0, 0, 3)


def test_bench_reads_measured_peaks_whatever_the_key_names():
    """MEASURED_PEAKS.json is written by the driver; bench.py must find the sustained and burst bf16 figures and the HBM
    copy figure under any reasonable naming, and must never crash on the file."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.parse_peaks({"bf16_tflops": 1632.4, "bf16_tflops_sustained": 1360.4, "hbm_gbs": 6557.8}) == (1360.4, 6557.8, 1632.4)
    assert b.parse_peaks({"bf16_tflops_sustained": 1386.8, "hbm_gbs": 6445.3}) == (1386.8, 6445.3, 1386.8)
    nested = {"hbm": {"copy_GBps": 6445.3}, "tensor": {"bf16_dense_tflops": {"burst": 1642.7, "sustained": 1386.8}},
              "sm_clock_mhz": {"median": 1342, "max": 1965}}
    assert b.parse_peaks(nested) == (1386.8, 6445.3, 1642.7)
    assert b.parse_peaks({"hbm_gb_s": 6445.3, "bf16_tflops_burst": 1642.7, "bf16_tflops": 1386.8}) == (1386.8, 6445.3, 1642.7)
    assert b.parse_peaks({"peaks": [{"name": "hbm_copy", "GB/s": 6445.3}, {"name": "bf16", "TFLOP/s": 1642.7}]}) == (1642.7, 6445.3, 1642.7)
    assert b.parse_peaks([]) == (None, None, None) and b.parse_peaks({"x": "y"}) == (None, None, None)
    pk = b.peaks()
    assert pk["burst"] >= pk["tensor"] > 0 and pk["hbm"] > 0
