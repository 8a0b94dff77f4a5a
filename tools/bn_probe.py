import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import oracle as O
from oracle import models_ref as MR
from nerf_keras_b200 import _lib
L_ = _lib.lib()
def run(Lay, H, skip, default_bn, B=96, Nc=16, Nf=32, seed=0, which="gb"):
    kw = dict(num_layers=Lay, hidden_dim=H, skip_layer=skip, lxyz=10, ldir=4)
    rng = np.random.default_rng(seed)
    wc, wf = O.init_weights(seed=1, bias_range=0.1, **kw), O.init_weights(seed=2, bias_range=0.1, **kw)
    bns = []
    for _ in range(2):
        bn = MR.init_bn(Lay, H)
        if not default_bn:
            for st in bn.values():
                n = st["gamma"].numel()
                g_ = torch.from_numpy(rng.uniform(0.7, 1.3, n).astype(np.float32)); b_ = torch.from_numpy(rng.uniform(-0.2, 0.2, n).astype(np.float32))
                if "g" in which: st["gamma"] = g_
                if "b" in which: st["beta"] = b_
        bns.append(bn)
    o, d = O.get_rays(12, 12, 15.0, torch.from_numpy(np.asarray(O.pose_spherical(25.0, -35.0, 4.0))))
    o, d = o.reshape(-1, 3)[:B].contiguous(), d.reshape(-1, 3)[:B].contiguous()
    t = O.generate_t_vals(2.0, 6.0, B, Nc, True, u=torch.from_numpy(rng.random(Nc, dtype=np.float32)))
    u = torch.from_numpy(rng.random((B, Nf), dtype=np.float32)); img = torch.from_numpy(rng.random((B, 3), dtype=np.float32))
    params = MR._params(wc) + MR._params(wf)
    for p in params: p.requires_grad_(True)
    b2 = [{r: {k: v.clone() for k, v in st.items()} for r, st in bn.items()} for bn in bns]
    rgbs = O.forward_pass(wc, wf, o, d, t, 10, 4, Nf, u, training=True, stop_grad_samples=True, num_layers=Lay, skip_layer=skip,
                          bn_coarse=b2[0], bn_fine=b2[1])[0]
    loss = MR.mse(img, rgbs[0]) + MR.mse(img, rgbs[1])
    g = torch.autograd.grad(loss, params)
    for p in params: p.requires_grad_(False)
    ref = np.concatenate([x.reshape(-1).numpy() for x in g])
    cfg = _lib.NerfConfig(Lay, H, skip, 10, 4, Nc, Nf, B, 0, 0, 5e-4, 1)
    n, nbn = int(L_.nerf_param_count(C.byref(cfg))), int(L_.nerf_bn_param_count(C.byref(cfg)))
    flat_bn = lambda bn: np.concatenate([np.concatenate([bn[r][k].numpy() for r in bn]) for k in ("gamma", "beta", "mean", "var")])
    P = torch.from_numpy(np.concatenate([O.flatten_weights(wc), O.flatten_weights(wf)])).cuda()
    BN = torch.from_numpy(np.concatenate([flat_bn(bns[0]), flat_bn(bns[1])]).astype(np.float32)).cuda()
    G, BG, Mx = torch.empty(2 * n, device="cuda"), torch.empty(4 * nbn, device="cuda"), torch.empty(3, device="cuda")
    ws = torch.empty(int(L_.nerf_bn_workspace_bytes(C.byref(cfg), B)), dtype=torch.uint8, device="cuda")
    io = [x.cuda().contiguous() for x in (img, o, d, t, u)]
    _lib.check(L_.nerf_bn_forward_backward(C.byref(cfg), P.data_ptr(), BN.data_ptr(), *[x.data_ptr() for x in io], B, G.data_ptr(),
                                           BG.data_ptr(), Mx.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "fb")
    got = G.cpu().numpy()
    off = 0; out = []
    for net in "cf":
        for role, fi, fo in MR.layer_shapes(**kw):
            k = fi * fo
            a, b = got[off:off + k], ref[off:off + k]
            if net == "c": out.append(f"{role}:{np.linalg.norm(a - b) / np.linalg.norm(b):.5f}")
            if net == "c" and role in ("d6", "d7") and which == "b":
                A2, B2 = a.reshape(fi, fo), b.reshape(fi, fo)
                ce = np.linalg.norm(A2 - B2, axis=0) / (np.linalg.norm(B2, axis=0) + 1e-30)
                re = np.linalg.norm(A2 - B2, axis=1) / (np.linalg.norm(B2, axis=1) + 1e-30)
                top = np.argsort(-ce)[:4]; topr = np.argsort(-re)[:4]
                beta = bns[0][role]["beta"].numpy()
                print(f"   {role}: worst output units {[(int(j), round(float(ce[j]), 4), round(float(beta[j]), 3), float(np.linalg.norm(B2[:, j]))) for j in top]}")
                print(f"   {role}: worst input rows {[(int(j), round(float(re[j]), 4), float(np.linalg.norm(B2[j]))) for j in topr]}  median col err {np.median(ce):.5f} median row err {np.median(re):.5f}")
            off += k + fo
    print(f"L={Lay} H={H} which={which} default_bn={default_bn} B={B}: coarse W rel err " + " ".join(out), flush=True)
run(8, 256, 4, False, which="b")
