#!/bin/bash
# Fetch the English C4 shards used for real-data runs (reference: scripts/pull-c4.sh). Needs git-lfs and network access.
set -e
GIT_LFS_SKIP_SMUDGE=1 git clone https://huggingface.co/datasets/allenai/c4
cd c4
git lfs pull --include "en/*"
