"""not-gpu: the C-ABI library loads and exports every symbol include/tcdiff_b200.h declares, the drop-in host
classes keep the reference contract, nothing silently falls back to the CPU, and the N>1 sharding logic works
(gloo, world_size 2)."""
import copy
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tcdiff_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "tcdiff_b200.h")).read()
    declared = set(re.findall(r"\b(tcd_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 24
    lib = _lib.lib()                                      # raises if the .so is missing
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.tcd_arch() == b"sm_100a" and lib.tcd_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (tcd_[a-z0-9_]+)", out))
    assert declared <= exported


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The build is sm_100a-native: tcgen05.mma -> UTCHMMA, TMA -> UTMALDG/UTMASTG, tcgen05.ld -> LDTM."""
    from tcdiff_b200 import _lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_argument_errors_are_reported_not_swallowed():
    from tcdiff_b200 import _lib
    lib = _lib.lib()
    rc = lib.tcd_cfg_ddim_step(0, 0, 0, 0, 0, 0, 0, 0, 0, 10, 150, 2.0, 1.0, 1.0, 1.0, 0.0, 0.0, 1, 0, 0)
    assert rc == -1 and b"151" in lib.tcd_last_error()
    with pytest.raises(_lib.TcdError):
        _lib.check(rc)
    assert lib.tcd_gemm(7, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0) == -1          # bad dtype / null pointers
    assert lib.tcd_cfg_ddim_step(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 151, 2.0, 1.0, 1.0, 1.0, 0.0, 0.0, 1, 0, 0) == 0  # empty input
    # block tails: a call that would produce nothing / misses required operands is rejected before any launch
    assert lib.tcd_film_residual_norm(1, 16, 0, 16, 1, 0, 0, 0.0, 0, 0, 0, 0, 0, 0.0, 0, 0, 0, 0, 8, 512, 4, 0) == -1
    assert b"x_out" in lib.tcd_last_error()
    assert lib.tcd_gemm_film_residual_norm(16, 512, 16, 512, 0, 8, 512, 16, 16, 0, 0, 0.0, 0, 0, 0, 16, 16, 1e-5, 0, 0, 0, 0,
                                           0, 4, 0) == -1                               # no output operand requested
    assert b"output" in lib.tcd_last_error()
    assert lib.tcd_gemm_film_residual_norm(16, 512, 16, 512, 0, 8, 512, 16, 16, 0, 0, 0.0, 0, 0, 0, 16, 16, 1e-5, 0, 16, 0, 0,
                                           0, 4, 0) == -1                               # rotary output without its tables
    assert lib.tcd_gemm_film_residual_norm(16, 512, 16, 512, 0, 8, 512, 16, 16, 0, 0, 0.0, 0, 0, 0, 16, 16, 1e-5, 0, 16, 16, 16,
                                           3, 4, 0) == -1                               # transposed tables shorter than a sample
    assert lib.tcd_gemm_frn_set_debug(0) == 0


def test_engine_defaults():
    """The library's compile-time tuning choices (csrc/tuning.cuh) are what the engine runs with: no environment variable
    switches a kernel variant anywhere in the product path (round-1 verdict), the dead feed-forward residual is skipped."""
    from tcdiff_b200 import engine, _lib
    assert engine.SKIP_DEAD_X is True
    lib = _lib.lib()
    assert lib.tcd_tuning(b"attn_2q") in (0, 1) and lib.tcd_tuning(b"fuse_tails") in range(8)
    assert lib.tcd_tuning(b"no_such_choice") == -1
    assert engine.fuse_tails() == lib.tcd_tuning(b"fuse_tails")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tcdiff_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".py")) and f != "build.py":
                src = open(os.path.join(dirpath, f)).read()
                assert "getenv" not in src, f
                assert not re.search(r"os\.environ[^\n]*TCD_", src), f


def test_no_cpu_fallback():
    import tcdiff_b200 as T
    cfg = synth.CONFIGS["tiny"]
    m = T.DanceDecoder(nfeats=151, seq_len=150, latent_dim=512, ff_size=cfg["ff_size"], num_layers=cfg["num_layers"],
                       num_heads=8, cond_feature_dim=cfg["cond_feature_dim"], required_dancer_num=cfg["dancers"]).eval()
    with pytest.raises(T.TcdError):
        m(torch.zeros(1, 300, 151), torch.zeros(1, 301, 13), torch.zeros(1).long())
    with pytest.raises(T.TcdError):
        T.ax_from_6v(torch.zeros(4, 6))
    with pytest.raises(T.TcdError):                       # train mode (bf16 tape, dropout 0.1): still CUDA only
        m.train()(torch.zeros(1, 300, 151), torch.zeros(1, 301, 13), torch.zeros(1).long())
    with pytest.raises(NotImplementedError):              # dropout > 0 exists on the bf16 tape only
        m.train().set_compute_dtype("fp32")(torch.zeros(1, 300, 151), torch.zeros(1, 301, 13), torch.zeros(1).long())
    m.set_compute_dtype("bf16")
    with pytest.raises(NotImplementedError):
        m.eval()(torch.zeros(1, 300, 151), torch.zeros(1, 301, 13), torch.zeros(1).long(), trj_dist=torch.zeros(1))
    # the product never imports the oracle
    src = "".join(open(os.path.join(ROOT, "tcdiff_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "tcdiff_b200"))
                  if f.endswith(".py"))
    assert "oracle" not in src
    # bench.py: only the CPU legs (cpu_baseline / --impl reference) may import the checker; the measured arm takes its
    # synthetic weights and inputs from tcdiff_b200/synth.py
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = [n for n in ast.walk(fn) if isinstance(n, (ast.Import, ast.ImportFrom))
                and "oracle" in (getattr(n, "module", None) or "") + " ".join(a.name for a in n.names)]
        assert not uses or fn.name in ("cpu_train_sample", "cpu_reference_sample"), fn.name
    assert not [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))
                and "oracle" in (getattr(n, "module", None) or "") + " ".join(a.name for a in n.names)]


def test_dropin_module_contract():
    import tcdiff_b200 as T
    cfg = synth.CONFIGS["tiny"]
    m = T.DanceDecoder(nfeats=151, seq_len=150, latent_dim=512, ff_size=cfg["ff_size"], num_layers=cfg["num_layers"],
                       num_heads=8, cond_feature_dim=cfg["cond_feature_dim"], required_dancer_num=cfg["dancers"])
    sd = synth.make_state_dict(cfg, 0)
    assert set(m.state_dict()) == set(sd)
    m.load_state_dict(sd, strict=True)
    m.load_state_dict({"module." + k: v for k, v in sd.items()}, strict=True)      # TCDiff.py:31-36
    dead = [k for k, _ in m.named_parameters() if "traj_Modulation" in k or "traj_embedding" in k or "embeddings_table" in k]
    assert len(dead) == cfg["num_layers"] * 15 + 5
    d = T.GaussianDiffusion(m, 150, 151, T.SMPLSkeleton(), schedule="cosine", n_timestep=1000, predict_epsilon=False,
                            loss_type="l2", use_p2=False, cond_drop_prob=0.25, guidance_weight=2)
    assert d.master_model is not m and set(d.master_model.state_dict()) == set(sd)
    assert d.master_model._cache.packed is None                 # derived cache does not survive deepcopy
    copy.deepcopy(d)
    # cache signature reacts to in-place optimizer-style updates and to EMA
    s0 = m._signature()
    with torch.no_grad():
        m.final_layer.bias.add_(1.0)
    assert m._signature() != s0
    with pytest.raises(T.TcdError):                             # EMA is a kernel too: no CPU path (GPU test: test_gpu_train)
        d.ema.update_model_average(d.master_model, d.model)
    with pytest.raises(NotImplementedError):
        T.DanceDecoder(nfeats=151, use_rotary=False)
    # the reference constructor's defaults (loss_type "l1", predict_epsilon=True) are on the supported path: the call gets as
    # far as the first kernel and fails there for lack of a GPU, not with NotImplementedError (GPU parity: test_gpu_model)
    dflt = T.GaussianDiffusion(m, 150, 151, None)
    assert dflt.predict_epsilon and dflt.loss_type == "l1"
    with pytest.raises(T.TcdError):
        dflt.p_sample_loop((1, 300, 151), torch.zeros(1, 301, 13))


def test_shard_bounds_cover_batch():
    from tcdiff_b200.dist import shard_bounds
    for B in (1, 2, 7, 64, 256):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from tcdiff_b200.dist import sharded_sample
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
B, L = 5, 6
g = torch.Generator().manual_seed(0)
cond = torch.randn(B, 3, 4, generator=g); x0 = torch.randn(B, L, 3, generator=g)
bank = [torch.randn(B, L, 151, generator=g) for _ in range(3)]
def sample_fn(shape, c, x_0=None, noise_bank=None):           # row-wise stand-in for ddim_sample (CPU)
    assert shape[0] == c.shape[0] == x_0.shape[0] == noise_bank[0].shape[0]
    out = sum(noise_bank) + c.sum((1, 2))[:, None, None]
    out[..., 4:6] = x_0[..., :2]
    return out
full = sharded_sample(sample_fn, (B, L, 151), cond, x_0=x0, noise_bank=bank)
ref = sample_fn((B, L, 151), cond, x_0=x0, noise_bank=bank)
assert full.shape == ref.shape and torch.equal(full, ref), "gathered result differs"
dist.destroy_process_group()
print("ok", sys.argv[3])
'''


def test_sharded_sampling_gloo_world2(tmp_path):
    """One process per rank, gloo backend: contiguous row shards + ONE all_gather reproduce the unsharded result
    (uneven split 3 + 2)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


_REDUCER_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from tcdiff_b200.dist import GradReducer
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2)
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(6, 40), torch.nn.ReLU(), torch.nn.Linear(40, 40), torch.nn.Tanh(),
                          torch.nn.Linear(40, 3))
dead = torch.nn.Parameter(torch.ones(5))                    # never used: its bucket must still complete
params = list(net.parameters()) + [dead]
offs, off = [], 0
for p in params:
    offs.append(off); off += (p.numel() + 7) // 8 * 8
arena = torch.zeros(off)
for p, o in zip(params, offs):
    p.grad = arena[o:o + p.numel()].view(p.shape)
red = None
for it in range(3):
    arena.zero_()
    if red is not None:
        red.reset()
    if it == 2:
        for p in net.parameters():
            p.grad = None                                   # cleared from outside: the hook must bring it home
    x = torch.randn(4, 6, generator=torch.Generator().manual_seed(10 * it + rank))
    net(x).square().sum().backward()
    if red is None:                                         # step 0: arena built after the first backward
        red = GradReducer(params, arena, offs, bucket_mb=100 * 4 / (1 << 20))
        assert len(red.buckets) >= 3, red.buckets
        red.finish(all_in_hooks=False)
    else:
        red.finish()
    for i, p in enumerate(list(net.parameters())):
        want = 0
        for r in range(2):
            xr = torch.randn(4, 6, generator=torch.Generator().manual_seed(10 * it + r))
            want = want + torch.autograd.grad(net(xr).square().sum(), p)[0]
        assert p.grad.data_ptr() == red.views[id(p)].data_ptr()
        assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-6), (it, i)
    assert float(dead.grad.abs().max()) == 0.0
assert red.launched > 0
dist.destroy_process_group()
print("ok", rank)
'''


def test_grad_reducer_gloo_world2(tmp_path):
    """Bucketed, hook-driven SUM all-reduce of the flat gradient arena (data-parallel training) over
    gloo with 2 ranks: arena == sum of the two ranks' gradients; unused parameters do not stall a bucket."""
    script = tmp_path / "reducer_worker.py"
    script.write_text(_REDUCER_WORKER)
    port = str(29950 + os.getpid() % 40)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_ctypes_signatures_match_header_arity():
    """Every SIGNATURES entry has as many argtypes as the header declaration has parameters (a stale binding would
    silently pass garbage through ctypes)."""
    from tcdiff_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "tcdiff_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for name, params in re.findall(r"\b(tcd_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert len(_lib.SIGNATURES[name]) == n, (name, len(_lib.SIGNATURES[name]), n)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/tcdiff_b200.h is the drop-in boundary: it must compile as strict C99 (no C++/torch types) and a plain C
    program must link against the shared library and call it."""
    import shutil
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    from tcdiff_b200 import _lib
    src = tmp_path / "abi.c"
    src.write_text('#include "tcdiff_b200.h"\n#include <string.h>\n'
                   'int main(void) { return (tcd_version() == 1 && strcmp(tcd_arch(), "sm_100a") == 0 && '
                   'tcd_gemm(7, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0) == TCD_ERR_INVALID && strlen(tcd_last_error()) > 0) ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "abi"
    r = subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-ltcdiff_sm100a",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_denoise_step_launch_census(monkeypatch):
    """The launch sequence of ONE sampler denoise step (cond + uncond, shared front) with the C-ABI calls recorded on
    CPU tensors, for the 8-layer geometry of BASELINE c2: the census must equal the `kernel_breakdown` launch counts of the
    measured bench line (profiles/r01_bench_n_rowstore.json: 69 GEMMs, 16 attention, 25 FiLM tails, 8 LayerNorm+rotary,
    2 scatter_rows), every GEMM / tail covers the rows it should, and the tails chain x -> plain/rot consistently."""
    from tcdiff_b200 import engine, ops
    from tcdiff_b200.diffusion import GaussianDiffusion
    cfg = dict(synth.CONFIGS["c2"])
    pw = engine.PackedWeights(synth.make_state_dict(cfg, 0), cfg, torch.bfloat16, torch.device("cpu"))
    D, dn, S, NL = cfg["latent_dim"], cfg["dancers"], cfg["seq_len"], cfg["num_layers"]
    assert NL == 8 and D == 512
    L, B = S * dn, 1
    R, NLD, Mm = 2 * B * L, NL * D, S + 2
    calls = []
    for name in ("gemm", "attention", "layernorm_rotary", "film_residual_norm", "gemm_film_residual_norm", "scatter_rows",
                 "convert_pad"):
        monkeypatch.setattr(ops, name, lambda *a, _n=name, **k: calls.append((_n, a, k)))
    monkeypatch.setattr(engine, "fuse_tails", lambda: 0)
    cpu = torch.device("cpu")
    ws = engine.Workspace(cpu)
    tab = dict(NLD=NLD, Mm=Mm, Kt=torch.empty(50, 2, NLD, dtype=torch.bfloat16), Vt=torch.empty(50, 2, NLD, dtype=torch.bfloat16),
               Kc=torch.empty(2 * B, Mm, NLD, dtype=torch.bfloat16), Vc=torch.empty(2 * B, Mm, NLD, dtype=torch.bfloat16),
               film_all=torch.empty(50, 2 * B, NL * 3 * 2 * D))
    x = torch.empty(B * L, 151)
    xpad = torch.empty(B * L, pw.in_w.shape[1], dtype=torch.bfloat16)
    out = torch.empty(R, 151)
    GaussianDiffusion._denoise_step(None, engine.Denoiser(pw), ws, tab, 3, x, xpad, B, out)
    census = {}
    for c in calls:
        census[c[0]] = census.get(c[0], 0) + 1
    assert census == {"scatter_rows": 2, "gemm": 69, "attention": 16, "film_residual_norm": 25, "layernorm_rotary": 8}
    # rows: layer 0's self-attention block runs on the B shared samples, everything after it on 2B
    gemm_rows = [c[2]["M"] for c in calls if c[0] == "gemm"]
    assert gemm_rows[:4] == [B * L, B * S, B * S, B * S]                 # input_projection + the 3 fusion linears
    assert gemm_rows[4:7] == [B * L] * 3 and set(gemm_rows[7:]) == {R}   # sa_qk, sa_v, sa_fc of layer 0; then 2B samples
    att = [c[1][12] for c in calls if c[0] == "attention"]               # `samples` argument
    assert att == [B] + [2 * B] * 15
    frn_rows = [c[1][15] for c in calls if c[0] == "film_residual_norm"]
    assert frn_rows == [B * L, B * L] + [R] * 23
    # the dead feed-forward residual is skipped by default: exactly one tail per layer has x_out = None
    assert sum(1 for c in calls if c[0] == "film_residual_norm" and c[1][2] is None) == (NL if engine.SKIP_DEAD_X else 0)
    # the head GEMM writes the (2B*L, 151) output with pitch 151
    last = [c for c in calls if c[0] == "gemm"][-1]
    assert last[1][4] is out and last[2]["N"] == 151 and last[2]["ldc"] == 151


def test_ddpm_guidance_weight_clipping_by_timestep():
    """model/diffusion.py:219-224: weight = min(w, 0) for t > T, min(w, 1) for t < 0.1 T, else w (host scalar per step)."""
    import tcdiff_b200 as T
    d = T.GaussianDiffusion(torch.nn.Linear(1, 1), 150, 151, T.SMPLSkeleton(), schedule="cosine", n_timestep=1000,
                            predict_epsilon=False, loss_type="l2", use_p2=False, cond_drop_prob=0.25, guidance_weight=2)
    assert [d._guidance_weight_at(i) for i in (0, 99, 100, 999, 1000, 1001)] == [1, 1, 2, 2, 2, 0]
    d.guidance_weight = 0.5
    assert [d._guidance_weight_at(i) for i in (0, 99, 100, 1001)] == [0.5, 0.5, 0.5, 0]
    d.guidance_weight = -1
    assert [d._guidance_weight_at(i) for i in (50, 500, 1001)] == [-1, -1, -1]
