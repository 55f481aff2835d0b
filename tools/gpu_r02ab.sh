#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gauss_u8_binomial" -c 1 -o gpurun_out/prof_gauss_binomial_k3 python tools/prof_ops.py gauss_u8_k3 > gpurun_out/ncu_gb.log 2>&1; tail -1 gpurun_out/ncu_gb.log | cut -c1-160
