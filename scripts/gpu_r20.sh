#!/bin/bash
mkdir -p gpurun_out
L3_MEGA_AHEAD=0 timeout 200 python scripts/mega_timeline.py llama3-8b 128 8 > gpurun_out/r20_tl_8b.log 2>&1
L3_MEGA_AHEAD=0 timeout 200 python scripts/mega_timeline.py llama3.2-1b 2048 8 > gpurun_out/r20_tl_1b.log 2>&1
L3_MEGA_AHEAD=0 timeout 200 python scripts/mega_timeline.py stories15M 8 8 float32 > gpurun_out/r20_tl_s.log 2>&1
head -75 gpurun_out/r20_tl_8b.log; head -40 gpurun_out/r20_tl_1b.log
