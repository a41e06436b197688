#!/usr/bin/env bash
# A/B of two builds of libvfuse.so on ONE box: build the candidate, copy it to scratch/libvfuse_new.so, check the baseline sources out again and rebuild,
# then run this through gpurun (boxes differ by up to 5 %, so only same-box alternating runs count)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
cp llm_quest_b200/libvfuse.so /tmp/old.so
for r in 1 2; do
cp /tmp/old.so llm_quest_b200/libvfuse.so; echo "== committed"; python tools/attn_bench.py 64 784 12 256 196 12 28 1764 12 16 6272 12 | cut -c15-90
cp scratch/libvfuse_new.so llm_quest_b200/libvfuse.so; echo "== next item decoded ahead"; python tools/attn_bench.py 64 784 12 256 196 12 28 1764 12 16 6272 12 | cut -c15-90
done
cp /tmp/old.so llm_quest_b200/libvfuse.so
