#!/bin/bash
for v in 0 1; do
for d in "512 512 256 31 31 41" "256 256 256 15 15 15"; do
FCB200_OTF_VARIANT=$v TOOL_SAVEMEMORY=1 python tools/pass_times.py 20 $d 2>&1 | grep '^{' | cut -c1-330
done; done
