#!/usr/bin/env python
"""train_lego.py --config config/*.json : entry point with the reference's CLI (train_lego.py:25-27) on the
B200 path.  Data are synthetic Lego-shaped views (the tiny-NeRF download is out of scope); launched under
torchrun it trains data-parallel like train_tpu_lego.py (one process per GPU, NCCL gradient all-reduce)."""
import argparse
import json
import os

import nerf_keras_b200 as nk
from nerf_keras_b200.config import load_config, model_kwargs
from nerf_keras_b200.dist import init_from_env
from nerf_keras_b200.synthetic import BatchedRayDataset, prepare_lego_data, prepare_fern_data


def main(dataset="lego"):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default=f"config/{dataset}_batch_debug.json")
    ap.add_argument("--steps-per-epoch", type=int, default=None)
    ap.add_argument("--views", type=int, default=10)
    ap.add_argument("--checkpoint-dir", type=str, default=None,
                    help="per-epoch validation render (PNG), weights and history JSON, as the reference's TrainCallback")
    ap.add_argument("--ndc", action="store_true",
                    help="fern only: forward-facing NDC rays, t in [0, 1] (extension; the reference's Fern rays are pinhole "
                         "with near / far from the bounds, fern_data_utils.py:489-496)")
    ap.add_argument("--data", type=str, default=None,
                    help="real data instead of the synthetic scene: tiny_nerf_data.npz or a Blender scene directory "
                         "(lego), an LLFF scene directory with poses_bounds.npy (fern)")
    args = ap.parse_args()
    conf = load_config(args.config)
    name = os.path.splitext(os.path.basename(args.config))[0]
    rank, local, world = init_from_env()
    nk.set_random_seed(42)
    if args.data is not None:
        from nerf_keras_b200 import real_data
        if dataset == "fern":
            train, val, (near, far), focal = real_data.prepare_fern_data(conf["HEIGHT"], conf["WIDTH"], datadir=args.data)
            if args.ndc:
                H_, W_ = conf["HEIGHT"], conf["WIDTH"]
                train = (train[0],) + tuple(nk.ndc_rays(H_, W_, focal, 1.0, train[1], train[2]))
                val = (val[0],) + tuple(nk.ndc_rays(H_, W_, focal, 1.0, val[1], val[2]))
                near, far = 0.0, 1.0
        elif os.path.isdir(args.data):
            train, val, (near, far), focal = real_data.prepare_blender_data(conf["HEIGHT"], conf["WIDTH"], args.data)
        else:
            train, val, (near, far), focal = real_data.prepare_lego_data(conf["HEIGHT"], conf["WIDTH"], npz_path=args.data)
    else:
        if args.ndc and dataset != "fern":
            raise SystemExit("--ndc applies to the forward-facing Fern scene only")
        if dataset == "lego":
            train, val, (near, far), focal = prepare_lego_data(conf["HEIGHT"], conf["WIDTH"], n_views=args.views)
        else:
            train, val, (near, far), focal = prepare_fern_data(conf["HEIGHT"], conf["WIDTH"], n_views=args.views, ndc=args.ndc)
    B, Nc, Nf = conf["BATCH_SIZE"], conf["NS_COARSE"], conf["NS_FINE"]
    train_ds = BatchedRayDataset(*train, Nc, B, near, far, shuffle=True, steps_per_epoch=args.steps_per_epoch,
                                 rank=rank, world=world)
    val_ds = BatchedRayDataset(*val, Nc, conf.get("TEST_BATCH_SIZE", B), near, far, shuffle=False,
                               steps_per_epoch=min(4, max(1, val[0].shape[0] // B)))
    coarse = nk.create_nerf_complete_model(**model_kwargs(conf))
    fine = nk.create_nerf_complete_model(**model_kwargs(conf))
    trainer = nk.NeRFTrainer(coarse, fine, B // world, Nc, Nf, conf["L_XYZ"], conf["L_DIR"])
    trainer.compile(optimizer=nk.Adam(learning_rate=conf["LEARNING_RATE"]), loss_fn=nk.MeanSquaredError())
    callbacks = []
    if args.checkpoint_dir and rank == 0:
        from nerf_keras_b200.callbacks import TrainCallback
        tag = f"l{conf['NUM_LAYERS']}_d{conf['HIDDEN_DIM']}_n{Nc + Nf}_ep{conf['EPOCHS']}"
        callbacks.append(TrainCallback(trainer, val[1], val[2], conf["HEIGHT"], conf["WIDTH"], near, far, Nc,
                                       args.checkpoint_dir, weight_name=f"nerf_{dataset}_{tag}.npz",
                                       history_name=f"history_{tag}.json"))
    history = trainer.fit(train_ds, validation_data=val_ds, epochs=conf["EPOCHS"], callbacks=callbacks,
                          verbose=int(rank == 0))
    if rank == 0:
        os.makedirs("models", exist_ok=True)
        trainer.save_weights(f"models/nerf_{dataset}_l{conf['NUM_LAYERS']}_d{conf['HIDDEN_DIM']}_n{Nc + Nf}_{name}.npz")
        with open(f"models/history_{name}.json", "w") as f:
            json.dump(history, f)
    if world > 1:
        from nerf_keras_b200.dist import shutdown
        shutdown(trainer)       # the step graphs hold NCCL kernels: release them before the process group goes away


if __name__ == "__main__":
    main("lego")
