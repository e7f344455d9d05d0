"""CPU-side tests (run with -m "not gpu"): C-ABI surface, drop-in module interfaces, host logic, data-parallel
gradient averaging over gloo (world_size 2). No CUDA compute."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "avt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avt_[a-z0-9_]+)\s*\(", src)))


def test_abi_library_exports_every_declared_symbol():
    from avt_b200 import _lib
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/avt_b200.h but not exported"
    bound = set(_lib.SIGNATURES) | set(_lib._SPECIAL)
    assert set(names) == bound, set(names) ^ bound
    assert _lib.lib().avt_abi_version() >= 1


def test_no_gpu_reports_error_not_fallback():
    from avt_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib().avt_check_device() == -3
    assert b"CUDA" in _lib.lib().avt_last_error() or b"device" in _lib.lib().avt_last_error()


def test_backbone_interface_matches_reference_timm_names():
    from avt_b200 import backbone
    from oracle import vit as o_vit
    ours = backbone.TIMMModel(1, "vit_base_patch16_224")
    ref = o_vit.create_model("vit_base_patch16_224")
    a = {k: tuple(v.shape) for k, v in ours.model.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert a == b
    assert hasattr(ours, "model")                       # train.init_from_model=[[backbone.model, ...]]
    assert backbone.create_model("vit_large_patch16_224").embed_dim == 1024
    with pytest.raises(RuntimeError):                   # no CPU path
        ours(torch.zeros(1, 3, 1, 224, 224))
    with pytest.raises(NotImplementedError):
        backbone.TIMMModel(1, "resnet50")


def test_avth_interface_matches_reference_names_and_errors():
    from avt_b200 import future_prediction as fp
    from oracle import avth as o_avth
    kw = dict(output_len=1, inter_dim=128, n_head=4, n_layer=3, return_past_too=True, avg_last_n=1)
    ours = fp.AVTh(64, future_pred_loss={"_target_": "torch.nn.MSELoss"}, future_pred_loss_wt=1.0, **kw)
    ref = o_avth.AVTh(64, future_pred_loss="mse", **kw)
    a = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert a == b
    assert ours.output_dim == 64 and isinstance(ours.future_pred_loss, torch.nn.MSELoss)
    assert ours.future_pred_loss.reduction == "none"
    assert tuple(ours.gpt_model.h[0].attn.c_attn.weight.shape) == (128, 384)   # HF Conv1D: [in, out]
    # BaseModel._initialize_weights only touches nn.Linear: encoder/decoder are, Conv1D containers are not
    assert isinstance(ours.encoder, torch.nn.Linear) and not isinstance(ours.gpt_model.h[0].attn.c_attn, torch.nn.Linear)
    with pytest.raises(RuntimeError):
        ours(torch.zeros(2, 5, 64), (2,))
    for bad in (dict(in_features=1), dict(in_features=64, assign_to_centroids="x"), dict(in_features=64, drop_last_n=1),
                dict(in_features=64, quantize_before_rollout=True)):
        with pytest.raises(NotImplementedError):
            fp.AVTh(**bad)


def test_checkpoint_key_compatibility():
    """timm ImageNet checkpoints (extra `head.*`) load into backbone.model with strict=False exactly as
    func/train.py:679-688 does, and head checkpoints written under transformers 4.2.2 (per-layer causal-mask buffers
    `attn.bias` / `attn.masked_bias`) load STRICTLY, also through a parent module's prefix."""
    import torch
    from avt_b200.backbone import TIMMModel
    from avt_b200.future_prediction import AVTh
    bb = TIMMModel(1, "vit_test_patch16_32")
    sd = {k: torch.randn_like(v) for k, v in bb.model.state_dict().items()}
    sd["head.weight"], sd["head.bias"] = torch.zeros(10, sd["norm.weight"].numel()), torch.zeros(10)
    res = bb.model.load_state_dict(sd, strict=False)
    assert res.missing_keys == [] and sorted(res.unexpected_keys) == ["head.bias", "head.weight"]
    assert torch.equal(bb.model.state_dict()["blocks.0.attn.qkv.weight"], sd["blocks.0.attn.qkv.weight"])
    head = AVTh(64, n_head=2, n_layer=2, inter_dim=64, n_positions=32)
    hsd = {k: torch.randn_like(v) for k, v in head.state_dict().items()}
    for i in range(2):
        hsd[f"gpt_model.h.{i}.attn.bias"] = torch.ones(1, 1, 32, 32)
        hsd[f"gpt_model.h.{i}.attn.masked_bias"] = torch.tensor(-1e4)
    head.load_state_dict(dict(hsd), strict=True)
    assert torch.equal(head.state_dict()["gpt_model.h.1.attn.c_attn.weight"], hsd["gpt_model.h.1.attn.c_attn.weight"])
    parent = torch.nn.Module()
    parent.future_predictor = AVTh(64, n_head=2, n_layer=2, inter_dim=64, n_positions=32)
    parent.load_state_dict({"future_predictor." + k: v for k, v in hsd.items()}, strict=True)


def test_split_k_heuristic():
    """Wave-aware split-K: units = tiles x split must not spill a few units into an extra round of the persistent grid."""
    from avt_b200.engine import _best_split, _split_k_for, small_m_split
    sk = _split_k_for(3072, 768, 15760, 256)             # 36 pair tiles on 74 CTA pairs -> split the 15760-row contraction
    assert sk >= 2 and (36 * sk) % 74 in (0, *range(60, 74))   # last round at least ~80 % full
    assert _split_k_for(2048, 8192, 80, 256) == 1        # AVT-h wgrad: K = 80 rows, nothing to split
    assert _split_k_for(768, 768, 15760, 256) <= 31
    # AVT-h weight-streaming GEMMs (M = 80): 48 / 32 / 64 / 32 tiles of 128 / 64 / 128 / 64 columns on 148 SMs
    from avt_b200.ops import small_m_block_n
    for N, K in [(6144, 2048), (2048, 2048), (8192, 2048), (2048, 8192)]:
        s = small_m_split(80, N, K)
        units = (N // small_m_block_n(N)) * s
        waves = -(-units // 148)
        # split factors are cluster sizes (1 / 2 / 4: the k-slabs of a tile reduce through distributed shared memory), one wave
        assert s in (1, 2, 4) and waves == 1 and units / 148 >= 0.6, (N, K, s)
    assert small_m_split(15760, 768, 768) == 1
    assert _best_split(10, 4, 148, 1, 6) == 1


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avt_b200.parallel import allreduce_mean_
    torch.manual_seed(rank)
    flat = [torch.full((1000,), float(rank + 1)), torch.arange(10, dtype=torch.float32) * (rank + 1)]
    allreduce_mean_(flat)
    q.put((rank, flat[0][0].item(), flat[1][3].item()))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_mean_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, 1.5, 4.5), (1, 1.5, 4.5)]        # mean over ranks (DDP semantics, func/train.py:771-778)


def test_bench_reference_arm_schema():
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--frames", "2"], capture_output=True, text=True, env=env, timeout=600)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_reads_measured_peaks_in_any_reasonable_schema(tmp_path, monkeypatch):
    """bench.peaks(): the driver-written MEASURED_PEAKS.json is preferred, the profiling recipe's numbers are the fallback."""
    import json
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.peaks() == (1400.0, 1590.0, 6650.0, "fallback")
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"bf16_tflops": 1639.1, "bf16_tflops_sustained": 1361.3,
                                                                "hbm_gbs": 6559.7}))
    assert bench.peaks() == (1361.3, 1639.1, 6559.7, "measured")
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"peaks": {"dense_bf16_TFLOPs_burst": 1639.1,
                                                                          "dense_bf16_TFLOPs_sustained": 1361.3,
                                                                          "HBM_copy_GBs": 6559.7}}))
    assert bench.peaks() == (1361.3, 1639.1, 6559.7, "measured")
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    assert bench.peaks()[3] == "fallback"


def test_finish_backward_interleaves_sgd_with_collective_waits():
    """FlatDataParallel.finish_backward(optimizer): this rank's AVT-h shard first (its reduce-scatter finished long ago, its
    update hides the tail of the collectives), then the backbone slices, then the torch-owned rest - each update only
    after ITS waits; the all-gather of the updated bf16 weights is left for the next step's begin_step()."""
    from avt_b200 import parallel
    from avt_b200.parallel import FlatDataParallel
    log = []

    class H:
        def __init__(self, name):
            self.name = name

        def wait(self):
            log.append("wait " + self.name)

    class Opt:
        def __init__(self, mods):
            self.mods = mods

        def sync_lr(self):
            log.append("lr")

        def step_flat(self, i):
            log.append(f"step_flat {i}")

        def step_flat_vectors(self, i):
            log.append(f"step_flat_vectors {i}")

        def step_flat_shard(self, i, shard, rng):
            log.append(f"step_flat_shard {i} {rng}")

        def step_other(self):
            log.append("step_other")

    class Pack:
        total, small_end = 1024, 512

    class Head:
        _pack = Pack()

    dp = FlatDataParallel.__new__(FlatDataParallel)
    dp.group, dp.comm_sms, dp.world = None, 0, 1
    dp.vit, dp.head = object(), Head()
    dp.other = []
    dp._other_handles, dp._other_early = [H("other")], set()
    dp._head_rs, dp._handles = ("shard", H("head")), [H("vit11"), H("vit0"), H("rest")]
    dp._head_small = [H("head vectors")]
    dp._shadow_shards_stale = dp._master_stale = False
    dp.finish_backward(Opt([dp.vit, dp.head]))
    assert log == ["lr", "wait head", "wait head vectors", "step_flat_vectors 1", "step_flat_shard 1 (512, 1024)", "wait vit11", "wait vit0", "wait rest", "step_flat 0",
                   "wait other", "step_other"]
    assert dp._handles == [] and dp._head_rs is None and dp._other_handles == []
    assert dp._shadow_shards_stale and dp._master_stale      # -> begin_step() all-gathers, state_dict() syncs the masters


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avt_b200.parallel import all_gather_shards_, reduce_scatter_mean, shard_range
    # gradients: rank r holds r+1 everywhere; the mean is 1.5; every rank keeps only its half
    g = torch.full((1024,), float(rank + 1)).to(torch.bfloat16)
    shard, h = reduce_scatter_mean(g)
    lo, hi = shard_range(1024)
    # "sharded optimizer": each rank updates only its shard of the weights, then the shards are all-gathered
    w = torch.zeros(1024)
    w[lo:hi] = -0.1 * shard.float() + rank
    all_gather_shards_(w)
    q.put((rank, (lo, hi), shard.dtype == torch.bfloat16, shard.float().unique().tolist(), w[0].item(), w[-1].item()))
    dist.destroy_process_group()


def test_reduce_scatter_update_all_gather_gloo_world2():
    """The sharded-optimizer exchange (reduce-scatter mean -> update of the local shard -> all-gather) over gloo, world 2."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7) % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res[0][:4] == (0, (0, 512), True, [1.5]) and res[1][:4] == (1, (512, 1024), True, [1.5])
    for r in res:     # every rank ends with both shards: rank 0's update in front, rank 1's behind
        assert abs(r[4] - (-0.15)) < 1e-6 and abs(r[5] - 0.85) < 1e-6


def test_launch_list_summary_splits_at_the_step_boundary(tmp_path):
    """tools/summarize_profiles.py launches_dram: the last step starts at its patchify launch (the first step carries one-time
    launches, so halving the list mis-splits); the mean DRAM bytes per ViT GEMM launch go into a JSON keyed by workload,
    which bench.py reports as roofline.traffic."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("summarize_profiles", os.path.join(root, "tools", "summarize_profiles.py"))
    sp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sp)
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"'
    rows, i = [hdr], 0

    def launch(name, ns, rd, wr, grid="(148, 1, 1)"):
        nonlocal i
        for metric, unit, val in (("dram__bytes_read.sum", "byte", rd), ("dram__bytes_write.sum", "byte", wr),
                                  ("gpu__time_duration.sum", "ns", ns)):
            rows.append(f'"{i}","1","python","h","{name}","1","7","(384, 1, 1)","{grid}","0","10.0","s","{metric}","{unit}","{val}"')
        i += 1

    gemm = "void avt::gemm_bf16_kernel<256, 2, 0, 0, 1>(CUtensorMap_st)"
    head_wgrad = "void avt::gemm_bf16_kernel<256, 2, 1, 1, 1>(CUtensorMap_st)"
    for step in range(2):
        if step == 0:
            for _ in range(5):
                launch("avt::cast_f32_bf16_kernel(const float *)", 1000, 10, 10)      # one-time launches of the first step
        launch("avt::patchify_kernel(const float *)", 17000, 48e6, 24e6)
        launch(gemm, 60000, 100e6, 20e6)
        launch(gemm, 40000, 60e6, 20e6)
        launch(head_wgrad, 15000, 1e6, 30e6)
        launch("avt::sgd_step_kernel<1>(float *)", 1000000, 3e9, 3e9)
    src, dst, js = tmp_path / "l.csv", tmp_path / "l.md", tmp_path / "t.json"
    src.write_text("\n".join(["==PROF== connected"] + rows) + "\n")
    sp.launches_dram(str(src), str(dst), str(js))
    text = dst.read_text()
    assert "5 launches" in text                                   # patchify + 2 GEMMs + head wgrad + ONE sgd launch
    assert "| `avt::sgd_step_kernel<1>` | 1 |" in text
    t = json.loads(js.read_text())["vit_base_patch16_224:10:8"]
    assert t["launches"] == 2 and abs(t["bytes_per_launch"] - 100e6) < 1     # (120 + 80) / 2 MB; the AVT-h weight gradient is left out
