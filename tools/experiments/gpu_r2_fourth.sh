#!/bin/bash
# round-2 captures of the bench frame (ncu_traffic.json source) + the surface-fetch kernel experiment
bash tools/gpu_r2_capture.sh default
bash tools/gpu_ab6.sh default lib_surf
cp gpurun_out/ab6.log gpurun_out/r2_surface_kernel_ab.log
