#!/bin/bash
mkdir -p gpurun_out; timeout 120 tools/microbench2 > gpurun_out/r17_microbench2.txt 2>&1; cat gpurun_out/r17_microbench2.txt
