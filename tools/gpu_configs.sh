#!/bin/bash
# config 2 (ReSTIR DI 720p) and config 5 (instanced field, ~50 M triangles) on one GPU (development aid)
mkdir -p gpurun_out
timeout 600 python tools/gpu_configs.py di > gpurun_out/config_di.log 2>&1
timeout 900 python tools/gpu_configs.py field ${1:-5} ${2:-28} > gpurun_out/config_field.log 2>&1
cat gpurun_out/config_di.log gpurun_out/config_field.log | cut -c1-1500
