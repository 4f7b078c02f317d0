#!/usr/bin/env python3
"""Multi-process leg of the row-strip tiling (BASELINE config 4): one process per GPU
under torchrun, halo rows over CUDA-IPC peer memory (NVLink), sums over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/tiled_check.py [--size 4096] [--steps 10] [--parity-size 256] [--oracle-size 1024]

1. parity: a small canvas is evaluated tiled over the N GPUs and, on rank 0, un-split on one GPU; loss,
   every trace value and the gathered gradient must agree (fp32 2e-5, fp16 1e-3 / 2e-3), then 3 L-BFGS
   steps must stay >= 60 dB (fp32) from the un-split trajectory.
2. oracle parity: the --oracle-size canvas tiled over the N GPUs against the CPU ORACLE (loss / trace / gradient).
3. timing: K L-BFGS iterations of the --size canvas, CUDA events, max over ranks.
Rank 0 prints one JSON line per part.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definition: images, weights, flop counts)


def psnr(a, b):
    a = np.clip(np.asarray(a, np.float64), 0, 255)
    b = np.clip(np.asarray(b, np.float64), 0, 255)
    mse = np.mean((a - b) ** 2)
    return 200.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=4096)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--parity-size', type=int, default=256)
    ap.add_argument('--precision', default='fp16')
    ap.add_argument('--oracle-size', type=int, default=1024,
                    help='canvas on which the strips are compared with the CPU ORACLE (0 = skip)')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from style_transfer2_b200.model import B200Model
    from style_transfer2_b200.tiled import TiledTransfer
    from style_transfer2_b200.worker import StyleTransfer

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- parity
    if args.parity_size:
        for precision, tol_s, tol_g, min_db in (('fp32', 2e-5, 2e-5, 60.0), ('fp16', 1e-3, 2e-3, 35.0)):
            model = B200Model(gpu=local, precision=precision)
            content, style, x0 = bench.load_images(args.parity_size)
            tt = TiledTransfer(model, x0.shape[0], x0.shape[1])
            tt.set_input(x0)
            tt.set_content(content)
            tt.set_style(style)
            tt.set_weights(bench.WEIGHTS, bench.PARAMS)
            loss, grads = tt.opfunc()
            grad = tt.gather(grads)
            grad = grad.cpu().numpy() if grad is not None else None
            tr = dict(tt.traces[-1].data)
            imgs = [tt.step()[0] for _ in range(3)]
            tt.check()
            out = {'part': 'parity', 'precision': precision, 'world': world, 'canvas': list(x0.shape[:2])}
            if rank == 0:
                ref = StyleTransfer(model)
                ref.set_input(x0)
                ref.set_content(content)
                ref.set_style(style)
                ref.set_weights(bench.WEIGHTS, bench.PARAMS)
                assert ref.start()
                _, g_ref = ref.opfunc(ref.input)
                tr_ref = ref.traces[-1].data
                g_ref = g_ref.cpu().numpy()
                worst = max(abs(tr[k] - v) / max(abs(v), 1e-30) for k, v in tr_ref.items() if k != 'time')
                gerr = float(np.linalg.norm(grad - g_ref) / np.linalg.norm(g_ref))
                dbs = [psnr(img, ref.step()[0]) for img in imgs]
                out.update(worst_trace_rel=worst, grad_rel=gerr, psnr_steps=dbs,
                           ok=bool(worst < tol_s and gerr < tol_g and min(dbs) > min_db))
                print(json.dumps(out), flush=True)
                assert out['ok'], out
            tt.close()
            barrier()

    # ---------------------------------------------------------------- strips vs the CPU oracle
    if args.oracle_size:
        model = B200Model(gpu=local, precision=args.precision)
        pj = bench.TiledJob(args.oracle_size, args.precision, model=model, prefill=0, want_first=True)
        pj.tt.check()
        if rank == 0:
            cpu_first = bench.oracle_first_eval(bench.oracle_job(args.oracle_size, full_net=False))
            par = bench.parity_of(pj.first, cpu_first)
            ok = par['loss_rel'] < 1e-3 and par['worst_loss_trace_rel'] < 1e-3 and par['grad_rel'] < 5e-2
            print(json.dumps(dict(par, part='oracle_parity', precision=args.precision, world=world,
                                  canvas=[args.oracle_size] * 2, ok=bool(ok))), flush=True)
            assert ok, par
        barrier()
        pj.close()

    # ---------------------------------------------------------------- timing
    model = B200Model(gpu=local, precision=args.precision)
    size = args.size
    content, style, x0 = bench.load_images(size)
    tt = TiledTransfer(model, size, size)
    tt.set_input(x0)
    tt.set_content(content)
    tt.set_style(style)
    tt.set_weights(bench.WEIGHTS, bench.PARAMS)
    for _ in range(max(args.warmup, 3)):
        tt.step(fetch=False)
    barrier()
    l0 = model.engine.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tt.step(fetch=False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tt.check()
    loss = float(tt.loss)
    if rank == 0:
        fl = 2 * bench.conv_flops(size, size)
        print(json.dumps({'part': 'timing', 'world': world, 'canvas': [size, size], 'precision': args.precision,
                          'steps': args.steps, 'ms_per_iteration': ms / args.steps,
                          'iterations_per_s': args.steps / (ms / 1000.0),
                          'conv_tflops_aggregate': fl / (ms / args.steps / 1000.0) / 1e12,
                          'launches_per_iteration_per_rank': (model.engine.launches() - l0) / args.steps,
                          'loss': loss}), flush=True)
    tt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
