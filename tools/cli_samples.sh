#!/bin/bash
# smvp-toolkit-cli --all-algs -n 1000 --json on the reference's sample matrices -> one JSON line per algorithm
# usage: tools/cli_samples.sh OUT.jsonl
set -e
OUT=${1:-gpurun_out/cli_samples_n1000.jsonl}
CLI=smvp-toolkit_b200/lib/smvp-toolkit-cli
TMP=$(mktemp -d)
: > "$OUT"
for m in ibm32 curtis54 pdp08-pg4 memplus pwt; do
  "$CLI" --all-algs -n 1000 --json -d "$TMP" tests/golden/sample-data/$m.mtx | grep '^{' | sed "s/^{/{\"matrix\": \"$m\", /" >> "$OUT"
done
rm -rf "$TMP"
cat "$OUT"
