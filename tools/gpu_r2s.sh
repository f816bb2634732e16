#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/time_config.py 480 640 16 10 96 10 2>&1 | grep sorted
timeout 300 python tools/time_config.py 480 640 8 10 96 10 2>&1 | grep sorted
NID_OPTS=task_px=16 timeout 300 python tools/time_config.py 480 640 16 10 96 10 2>&1 | grep sorted
