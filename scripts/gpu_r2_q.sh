#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_torchrun.py -m gpu -q -x -s -p no:cacheprovider --tb=long 2>&1 | tail -40 | cut -c1-1500
