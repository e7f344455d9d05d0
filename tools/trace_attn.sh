#!/bin/bash
# Per-phase timeline of CTA 0 of the attention kernels: rebuild with -DAVT_ATTN_TRACE on the GPU box, run once, restore.
cp avt_b200/libavt_b200.so /tmp/libavt_b200.so.keep
AVT_EXTRA_NVCC_FLAGS=-DAVT_ATTN_TRACE python -m avt_b200.build --force > /dev/null
python tools/profile_attn.py "$@"
cp /tmp/libavt_b200.so.keep avt_b200/libavt_b200.so
