#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
bash tools/gpu_var.sh base "$@"
