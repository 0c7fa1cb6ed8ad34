#!/bin/bash
python tools/gpu_cfg5.py 2>&1 | grep config5
for v in variants/libiactrace_b200_*.so; do [ -f "$v" ] && echo "$v" && IACTRACE_B200_LIB=$PWD/$v python tools/gpu_cfg5.py 2>&1 | grep config5; done
