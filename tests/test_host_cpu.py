"""CPU: host-side logic, config schema, C-ABI library loads and exports every declared symbol,
world_size-2 gloo test of the data-parallel plumbing.  No compute calls (no GPU here)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols(names=("nerf_b200.h", "nerf_b200_debug.h")):
    """Every function the public header (the drop-in boundary) and the debug header (tests / tools) declare."""
    out = set()
    for name in names:
        text = open(os.path.join(ROOT, "include", name)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        out |= set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", text))
    return sorted(out)


def test_library_exports_every_header_symbol():
    from nerf_keras_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libnerf_b200.so not built: run __graft_entry__.build()"
    handle = C.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    public = _header_symbols(("nerf_b200.h",))
    assert not [s for s in public if "selftest" in s or "debug" in s], "debug / self-test exports belong in nerf_b200_debug.h"
    for s in syms:
        assert hasattr(handle, s), f"missing export {s}"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in _lib.SIGNATURES"


def test_param_count_and_loud_failure_without_gpu():
    from nerf_keras_b200 import _lib
    L = _lib.lib()
    cfg = _lib.NerfConfig(8, 256, 4, 10, 4, 64, 128, 4096, 0, 0, 5e-4, 1)
    assert L.nerf_param_count(C.byref(cfg)) == 595844
    bad = _lib.NerfConfig(8, 256, 4, 10, 4, 64, 128, 4096, 1, 0, 5e-4, 1)  # BATCH_NORM=true
    assert L.nerf_param_count(C.byref(bad)) == -1
    assert b"BATCH_NORM" in L.nerf_last_error()
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert L.nerf_create(C.byref(cfg), C.byref(h)) != 0
        assert b"no CPU fallback" in L.nerf_last_error()
        import nerf_keras_b200 as nk
        with pytest.raises(RuntimeError):
            nk.get_rays(4, 4, 1.0, np.eye(4, dtype=np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nerf_keras_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_models_host_logic():
    import nerf_keras_b200 as nk
    from nerf_keras_b200.models import layer_shapes
    nk.set_random_seed(42)
    m = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=False)
    assert m.count_params() == 595844
    shapes = dict((r, (i, o)) for r, i, o in layer_shapes(8, 256, 4, 10, 4))
    assert shapes["d5"] == (319, 256) and shapes["ddir"] == (283, 128)
    w = m.get_weights()
    assert w["d5"]["W"].shape == (319, 256) and float(np.abs(w["d0"]["b"]).max()) == 0.0
    lim = np.sqrt(6.0 / (63 + 256))
    assert np.abs(w["d0"]["W"]).max() <= lim
    w["rgb"]["b"][:] = [1, 2, 3]
    m.set_weights(w)
    assert np.array_equal(m.get_weights()["rgb"]["b"], [1, 2, 3])
    # BATCH_NORM=true: inference folds gamma / sqrt(var + eps) into the Dense weights on the host
    mb = nk.create_nerf_complete_model(8, 256, 4, 10, 4, bn=True)
    assert sorted(mb.get_bn_params()) == sorted([f"d{i}" for i in range(8)] + ["ddir"])
    bn = mb.get_bn_params()
    bn["d0"]["gamma"][:] = 2.0; bn["d0"]["var"][:] = 4.0 - 1e-3; bn["d0"]["mean"][:] = 1.0; bn["d0"]["beta"][:] = 0.5
    mb.set_bn_params(bn)
    raw, dev = mb.get_weights()["d0"], mb.device_blob()
    np.testing.assert_allclose(dev[:63 * 256].reshape(63, 256), raw["W"], rtol=1e-6)           # s = 2 / sqrt(4) = 1
    np.testing.assert_allclose(dev[63 * 256:63 * 256 + 256], raw["b"] - 1.0 + 0.5, atol=1e-6)    # (b - mean) s + beta
    with pytest.raises(TypeError):
        nk.NeRFTrainer(object(), m, 8, 4, 8, 10, 4)
    oracle_like = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    tr = nk.NeRFTrainer(m, oracle_like, 8, 4, 8, 10, 4)
    assert [x.name for x in tr.metrics] == ["loss", "psnr"]


def test_pose_spherical_matches_reference_formula():
    import nerf_keras_b200 as nk
    c2w = nk.pose_spherical(30.0, -30.0, 4.0)
    assert c2w.shape == (4, 4) and c2w.dtype == np.float32
    assert abs(np.linalg.norm(c2w[:3, 3]) - 4.0) < 1e-5
    np.testing.assert_allclose(c2w[:3, :3] @ c2w[:3, :3].T, np.eye(3), atol=1e-6)


def test_config_schema_roundtrip():
    from nerf_keras_b200.config import load_config, REQUIRED_KEYS
    cfgdir = os.path.join(ROOT, "config")
    names = sorted(f for f in os.listdir(cfgdir) if f.endswith(".json"))
    assert "lego_batch_h256.json" in names and "fern_batch_h256_tpu.json" in names
    for n in names:
        conf = load_config(os.path.join(cfgdir, n))
        for k in REQUIRED_KEYS:
            assert k in conf
    with pytest.raises(KeyError):
        load_config(os.path.join(cfgdir, "lego_batch_h256.json"), override={"__drop__": "NS_FINE"})


def test_shard_range_partitions():
    from nerf_keras_b200.dist import shard_range
    for n, w in [(4096, 8), (10, 3), (7, 8), (0, 2)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        sizes = [e - s for s, e in spans]
        assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from nerf_keras_b200.dist import init_from_env, allreduce_sum_, shard_range, gather_rows
rank, local, world = init_from_env("gloo")
assert world == 2
g = torch.full((1000,), float(rank + 1))
w = allreduce_sum_(g)
assert w == 2 and torch.allclose(g, torch.full((1000,), 3.0)), g[:4]
# DP semantics: mean of per-shard mean-gradients == gradient of the global mean (equal shards)
full = torch.arange(64, dtype=torch.float32)
s, e = shard_range(64, rank, world)
local_grad = full[s:e].mean().reshape(1)
allreduce_sum_(local_grad)
assert abs(float(local_grad) / world - float(full.mean())) < 1e-6
rows = torch.full((3, 2), float(rank))
out = gather_rows(rows)
if rank == 0:
    assert out.shape == (6, 2) and float(out[3:].min()) == 1.0
else:
    assert out is None
dist.barrier()
sys.stdout.write("rank%d_ok\n" % rank); sys.stdout.flush()
'''


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script), ROOT],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank0_ok" in r.stdout and "rank1_ok" in r.stdout


def test_keras_bridge_role_table_matches_layer_shapes():
    """tools/keras_weights_bridge.py (run on the reference side) and the host mirror agree on roles and (in, out) shapes."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("keras_weights_bridge", os.path.join(ROOT, "tools", "keras_weights_bridge.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from nerf_keras_b200.models import layer_shapes
    for conf in ({"NUM_LAYERS": 8, "HIDDEN_DIM": 256, "SKIP_LAYER": 4, "L_XYZ": 10, "L_DIR": 4},
                 {"NUM_LAYERS": 4, "HIDDEN_DIM": 64, "SKIP_LAYER": 2, "L_XYZ": 6, "L_DIR": 2}):
        roles, shapes = mod.expected_shapes(conf)
        want = layer_shapes(conf["NUM_LAYERS"], conf["HIDDEN_DIM"], conf["SKIP_LAYER"], conf["L_XYZ"], conf["L_DIR"])
        assert roles == [r for r, _, _ in want]
        assert shapes == [(fi, fo) for _, fi, fo in want]


def test_callback_image_scaling_matches_array_to_img():
    """keras.utils.array_to_img(scale=True): (x - min) / (max - min) * 255, constant images stay 0 (train_lego.py:216-220)."""
    from nerf_keras_b200.callbacks import array_to_uint8
    x = np.array([[0.25, 0.5], [0.75, 1.25]], dtype=np.float32)
    np.testing.assert_array_equal(array_to_uint8(x), np.array([[0, 63], [127, 255]], dtype=np.uint8))
    assert array_to_uint8(np.full((2, 2), 3.0, np.float32)).max() == 0


def test_reference_literal_configs_load_unchanged():
    """config/ref_*.json hold the LITERAL values of the reference's six config files (the shipped config/*.json without the
    prefix are the benchmark-shaped variants BASELINE.json names).  They load through the same schema check, build the
    model kwargs, and -- in the container that has the reference tree -- equal the reference's files value for value."""
    import glob
    import json
    from nerf_keras_b200.config import load_config, model_kwargs
    files = sorted(glob.glob(os.path.join(ROOT, "config", "ref_*.json")))
    assert len(files) == 6
    for f in files:
        conf = load_config(f)
        kw = model_kwargs(conf)
        assert kw["num_layers"] == 8 and kw["hidden_dim"] == 256 and kw["skip_layer"] == 4 and kw["lxyz"] == 10 and kw["ldir"] == 4
        assert isinstance(conf["BATCH_NORM"], bool) and conf["LEARNING_RATE"] == 0.0005
        ref = os.path.join("/root/reference/config", os.path.basename(f)[len("ref_"):])
        if os.path.exists(ref):
            assert json.load(open(ref)) == conf
    # the reference's Fern scripts need TEST_BATCH_SIZE, which fern_batch_h256.json lacks (train_fern.py:38 raises KeyError)
    assert "TEST_BATCH_SIZE" not in load_config(os.path.join(ROOT, "config", "ref_fern_batch_h256.json"))
