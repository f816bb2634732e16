#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for o in "" "lm_speculate=0" "lm_speculate=2" "lm_speculate=4"; do NID_OPTS=$o timeout 300 python tools/time_single.py 2>&1 | tr '\n' ' '; echo; done
NID_OPTS= timeout 300 python tools/time_single.py 480 640 16 10 2>&1 | tr '\n' ' '; echo
