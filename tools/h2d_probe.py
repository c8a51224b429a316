#!/usr/bin/env python
"""Concurrent pinned host-to-device bandwidth of all ranks for different copy sizes / stream counts (development).

    torchrun --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/h2d_probe.py

What the from-stride-8 e2e leg needs to know: its loader batches are 5 MB each; are many small copies as fast as one big one
when every GPU of the node copies at the same time?  Prints one JSON line (aggregate GB/s over all ranks).
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    total = 640 << 20
    host = torch.empty(total, dtype=torch.uint8).pin_memory()
    host.fill_(3)
    small = torch.empty(20 << 20, dtype=torch.uint8).pin_memory()      # a 20 MB source that is cycled through (stays in the LLC)
    dst = torch.empty(total, dtype=torch.uint8, device=dev)
    res = {}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(name, chunk, n_streams, events, src):
        streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        main_stream = torch.cuda.current_stream(dev)
        n = total // chunk
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            for k in range(n):
                st = streams[k % n_streams]
                so = (k * chunk) % src.numel()
                with torch.cuda.stream(st):
                    dst[k * chunk:(k + 1) * chunk].copy_(src[so:so + chunk], non_blocking=True)
                    if events:
                        ev = torch.cuda.Event()
                        ev.record(st)
                        main_stream.wait_event(ev)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        t = torch.tensor([total / dt / 1e9], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        res[name] = round(float(t[0]), 1)

    run('one 640 MB copy', total, 1, False, host)
    run('40 MB copies, 1 stream', 40 << 20, 1, False, host)
    run('5 MB copies, 1 stream', 5 << 20, 1, False, host)
    run('5 MB copies, 1 stream, event per copy', 5 << 20, 1, True, host)
    run('5 MB copies, 2 streams, event per copy', 5 << 20, 2, True, host)
    run('5 MB copies, 4 streams, event per copy', 5 << 20, 4, True, host)
    run('5 MB copies from a cycled 20 MB source, 1 stream', 5 << 20, 1, True, small)
    run('5 MB copies from a cycled 20 MB source, 2 streams', 5 << 20, 2, True, small)
    run('2.5 MB copies, 2 streams', 2560 << 10, 2, True, host)

    # the other direction, and both at once (the e2e pipeline copies a window's results out while the next windows come in)
    back = torch.empty(160 << 20, dtype=torch.uint8).pin_memory()
    dsrc = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
    out_stream, in_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def duplex(name, with_h2d, with_d2h, d2h_chunk):
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            n = total // (5 << 20)
            for k in range(n):
                if with_h2d:
                    with torch.cuda.stream(in_stream):
                        dst[k * (5 << 20):(k + 1) * (5 << 20)].copy_(host[k * (5 << 20):(k + 1) * (5 << 20)], non_blocking=True)
                if with_d2h and k % 4 == 0:
                    o = (k // 4 * d2h_chunk) % (back.numel() - d2h_chunk + 1)
                    with torch.cuda.stream(out_stream):
                        back[o:o + d2h_chunk].copy_(dsrc[o:o + d2h_chunk], non_blocking=True)
            in_stream.synchronize()
            t_in = time.perf_counter() - t0
            torch.cuda.synchronize()
            t_all = time.perf_counter() - t0
        h2d = total / t_in / 1e9 if with_h2d else 0.0
        d2h = (n // 4) * d2h_chunk / t_all / 1e9 if with_d2h else 0.0
        t = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        res[name] = {'h2d': round(float(t[0]), 1), 'd2h': round(float(t[1]), 1)}

    duplex('d2h alone, 5 MB copies', False, True, 5 << 20)
    duplex('h2d 5 MB copies + d2h 5 MB per 20 MB in', True, True, 5 << 20)
    duplex('h2d 5 MB copies + d2h 2.5 MB per 20 MB in', True, True, 2560 << 10)
    if rank == 0:
        print(json.dumps({'n_gpus': world, 'aggregate_gb_per_s': res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
