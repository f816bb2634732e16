#!/bin/bash
for o in "" "task_px=16" "task_px=24" "task_px=20"; do NID_OPTS=$o timeout 300 python tools/time_single.py 2>&1 | tr '\n' ' '; echo; done
