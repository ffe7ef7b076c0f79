#!/bin/bash
# development helper: GPU tests + 1184-frame timing + per-phase cycle profile, one gpurun call.  usage: quick_gpu.sh <tag> [pytest-args]
tag=$1; shift
sel=${*:-tests -m gpu -x -q}
tools/gpu_retry.sh --timeout 700 -- "python -m pytest $sel > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log; python bench.py --frames 1184 --steps 3 --warmup 2 --no-legs --no-cpu-baseline > gpurun_out/${tag}_bench.json 2>gpurun_out/${tag}_bench.err; python -c \"
import json;d=json.loads([l for l in open('gpurun_out/${tag}_bench.json') if l.startswith('{')][-1]);print('step ms',d['ms_per_step'],'kernel ms',d['roofline']['kernel_ms_per_launch'],'nl ms',d['roofline']['neighbour_build_ms_per_chunk'],'warm',d['warm']['ms_per_step'])
\"; if [ -f mythos_b200/libmythos_b200_prof.so ]; then MYTHOS_B200_LIB=/root/repo/mythos_b200/libmythos_b200_prof.so python bench.py --frames 296 --steps 1 --warmup 1 --no-legs --no-cpu-baseline 2>&1 | grep 'frame-kernel' | tail -2; fi" 2>&1 | tail -9
