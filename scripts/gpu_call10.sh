#!/bin/bash
mkdir -p gpurun_out
./build/ifetch_bench > gpurun_out/r2_ifetch_roofline.jsonl 2>&1
cat gpurun_out/r2_ifetch_roofline.jsonl
