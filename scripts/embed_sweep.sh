#!/usr/bin/env bash
for lib in semtools_b200/lib/variants/libstb_embed_*.so; do
  echo -n "$(basename $lib) "; STB_LIB_PATH=$PWD/$lib timeout 120 python scripts/embed_probe.py 2>&1 | grep '^{' | tail -1
done
