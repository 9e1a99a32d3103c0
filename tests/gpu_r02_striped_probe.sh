#!/bin/bash
mkdir -p gpurun_out
python tests/striped_probe.py 2>&1 | tail -3 | tee gpurun_out/r02_striped_probe.jsonl
