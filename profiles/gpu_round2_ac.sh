#!/bin/bash
mkdir -p gpurun_out
timeout 900 python profiles/prof_rows_r02.py > gpurun_out/rows_r02.json 2> gpurun_out/rows_r02.err
tail -3 gpurun_out/rows_r02.err; cat gpurun_out/rows_r02.json
