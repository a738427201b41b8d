#!/bin/bash
# Builds the host layer + the player's MP4 reader with AddressSanitizer / UBSan and runs the mutation fuzzer over the seed
# corpus:  tests/fuzz/run.sh [iterations per seed] [rng seed] [stub]
#   default: linked against the real engine library (no GPU here: configure fails once everything is parsed);
#   "stub":  tests/fuzz/engine_stub.c stands in for libiamf_b200.so, so that the decode path (temporal units, parameter
#            blocks, codec glue, time lines) runs too.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
IT=${1:-20000}; SEED=${2:-1}; MODE=${3:-real}
OUT=${FUZZ_DIR:-/tmp/iamfb_fuzz}
mkdir -p $OUT
python $ROOT/tests/fuzz/make_corpus.py $OUT/corpus
CODECS=""
[ -f /root/reference/dep_codecs/lib/libopus.a ] && CODECS="$CODECS -DIH_HAVE_OPUS /root/reference/dep_codecs/lib/libopus.a"
[ -f /root/reference/dep_codecs/lib/libFLAC.a ] && CODECS="$CODECS -DIH_HAVE_FLAC /root/reference/dep_codecs/lib/libFLAC.a"
if [ "$MODE" = stub ]; then ENGINE="$ROOT/tests/fuzz/engine_stub.c"; else ENGINE="-L$ROOT/iac_b200 -l:libiamf_b200.so -Wl,-rpath,$ROOT/iac_b200"; fi
gcc -std=gnu99 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -I$ROOT/include -I$ROOT/iac_b200/host -I$ROOT/iac_b200/player \
  -o $OUT/fuzz_host $ROOT/tests/fuzz/fuzz_host.c $ROOT/iac_b200/host/*.c $ROOT/iac_b200/player/iamfb_mp4.c $CODECS $ENGINE -lm -lpthread
ASAN_OPTIONS=detect_leaks=1:abort_on_error=1:protect_shadow_gap=0 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 \
  $OUT/fuzz_host $IT $SEED $OUT/corpus/*.bin $OUT/corpus/*.mp4
