#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "replay_the_reference_tape" 2>&1 | tail -3
OBS=torch timeout 300 python scripts/gpu_vecenv_profile.py 2>&1 | head -45 | cut -c1-160
