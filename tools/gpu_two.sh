#!/bin/bash
for cfg in "BVG_CONV_EPIW=8" "BVG_CONV_EPIW=4"; do echo "== $cfg"; env $cfg timeout 300 python tools/two_stream.py 2>&1 | tail -5; done
