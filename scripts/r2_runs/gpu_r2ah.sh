#!/bin/bash
for w in chain100 random16; do CBX_LIB=marlon_b200/libcbx_trace.so WORKLOAD=$w timeout 300 python scripts/r2_runs/trace_fine.py 2>&1 | tail -42; done
