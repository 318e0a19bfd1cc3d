#!/bin/bash
COORDX_PRECISION=fp32 timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -v Warn | tail -n 22
