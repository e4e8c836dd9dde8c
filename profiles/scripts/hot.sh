#!/bin/bash
# usage: profiles/scripts/hot.sh <report> <kernel regex> [top]
ncu -i "$1" --page source --print-source cuda,sass --csv --kernel-name regex:"$2" 2>/dev/null > /tmp/src_hot.csv
python profiles/scripts/ncu_lines.py /tmp/src_hot.csv ${3:-45}
