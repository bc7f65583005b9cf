#!/bin/bash
# One command for a maintainer WITH a Rust toolchain (this repository was developed without one): builds the
# real `awry` 0.3.1, runs FmIndex::new / save / load / parallel_count / parallel_locate on the repo's seeded
# FASTAs with both crates and diffs outputs and `.awry` files (rust/awry-b200/tests/against_reference.rs).
#   scripts/pin_against_reference.sh [path/to/awry/checkout]
# Needs: cargo, a CUDA 12.9 toolchain and a B200 (the library has no CPU fallback).
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
command -v cargo >/dev/null || { echo "cargo not found: this check needs a Rust toolchain" >&2; exit 2; }
make -C "$ROOT/awry_b200/csrc"
export AWRY_B200_LIB_DIR="$ROOT/awry_b200"
export LD_LIBRARY_PATH="$ROOT/awry_b200:${LD_LIBRARY_PATH:-}"
cd "$ROOT/rust/awry-b200"
# `awry` enters as a dev-dependency only here, so that the crate otherwise builds without the registry
if ! grep -q '^\[dev-dependencies\]' Cargo.toml; then
  if [ $# -ge 1 ]; then
    printf '\n[dev-dependencies]\nawry = { path = "%s" }\n\n[features]\nagainst-reference = []\n' "$1" >> Cargo.toml
  else
    printf '\n[dev-dependencies]\nawry = "=0.3.1"\n\n[features]\nagainst-reference = []\n' >> Cargo.toml
  fi
fi
cargo test --release --features against-reference --test against_reference -- --test-threads 1 --nocapture
