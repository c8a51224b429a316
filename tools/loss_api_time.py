#!/usr/bin/env python
"""BASELINE configs[2] through bench.py's own loss leg (tools/bench_extra.loss_leg): kernel calls and `compute_loss` + `backward`
through the segmentor, without the rest of the bench.  Development tool."""
import sys, os, json, torch
sys.path.insert(0, os.getcwd())
from tools.bench_extra import loss_leg
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
out = loss_leg(torch.device('cuda'), peak)
print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != 'roofline'}) for k, v in out.items()}, indent=1))
