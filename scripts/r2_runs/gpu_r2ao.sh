#!/bin/bash
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python scripts/bench_line.py final | head -3
