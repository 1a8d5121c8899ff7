#!/bin/sh
# Produce reference-held vectors with the reference's own, unmodified code (see README.md).
#   usage: dump_vectors.sh <checkout of Liam-Eagen/BulletproofsPP> <output directory>
set -eu
REF=$(cd "$1" && pwd)
OUT=$(mkdir -p "$2" && cd "$2" && pwd)
HERE=$(cd "$(dirname "$0")" && pwd)
cd "$REF"
stack build
for ex in 64bit rec_test bin_test; do
    mkdir -p "$OUT/$ex"
    # prove writes commits.bin / proof.bin; --write-points 8 also writes points.bin (app/Main.hs:259-263)
    (cd "$OUT/$ex" && stack --stack-yaml "$REF/stack.yaml" exec BulletproofsPP-exe -- prove \
        "$REF/examples/$ex/schema.json" "$REF/examples/$ex/witness.json" commits.bin proof.bin --write-points 8 \
        > stdout.txt)
    # the stock verifier on its own files (records whether the file round trip works for this example)
    (cd "$OUT/$ex" && stack --stack-yaml "$REF/stack.yaml" exec BulletproofsPP-exe -- verify \
        "$REF/examples/$ex/schema.json" commits.bin proof.bin >> stdout.txt 2>&1 || true)
done
# generators, `show` format and the first challenges straight from the reference's definitions
stack ghci BulletproofsPP:exe:BulletproofsPP-exe --ghci-options "-ghci-script $HERE/dump_vectors.ghci" \
    < /dev/null > "$OUT/ghci.txt" 2>&1 || true
echo "wrote $OUT; now run: python $HERE/vectors_to_json.py $OUT"
