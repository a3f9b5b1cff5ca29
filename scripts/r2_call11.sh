#!/bin/bash
mkdir -p gpurun_out
echo "== 1b, context 2048"; timeout 300 python scripts/mega_timeline.py llama3.2-1b 2048 8 > gpurun_out/r2c11_mega_1b_2048.txt 2>&1; sed -n 20,40p gpurun_out/r2c11_mega_1b_2048.txt
echo "== 1b, context 128"; timeout 300 python scripts/mega_timeline.py llama3.2-1b 128 8 > gpurun_out/r2c11_mega_1b_128.txt 2>&1; sed -n 20,40p gpurun_out/r2c11_mega_1b_128.txt
echo "== 8b, context 128"; timeout 300 python scripts/mega_timeline.py llama3-8b 128 8 > gpurun_out/r2c11_mega_8b_128.txt 2>&1; sed -n 20,40p gpurun_out/r2c11_mega_8b_128.txt
