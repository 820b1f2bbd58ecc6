#!/bin/sh
# Does the reference encoder's output depend on indeterminate (uninitialised) automatic variables?  Builds two DIAGNOSTIC copies of
# the reference into /tmp (never into the repo) that differ only in what gcc writes into uninitialised locals
# (-ftrivial-auto-var-init=zero / =pattern) and encodes the same clip with both.  Different streams = yes.
# usage: tools/ref_uninit_probe.sh [WxHxN]   (needs /root/reference)
set -e
CLIP=${1:-1280x720x1}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REFSRC=/root/reference/src/homer_lib
for MODE in zero pattern; do
  D=/tmp/hb_refdiag_$MODE
  mkdir -p $D/obj
  for f in $REFSRC/*.c; do
    o=$D/obj/$(basename $f .c).o
    [ -f $o ] || gcc -O3 -w -fmessage-length=0 -msse -msse2 -mssse3 -msse4 -msse4.1 -msse4.2 -fPIC -ftrivial-auto-var-init=$MODE -c -o $o $f
  done
  gcc -shared -o $D/libhomer_ref.so $D/obj/*.o -lpthread -lm
  gcc -O2 -w -fPIC -shared -msse4.2 -I$REFSRC -o $D/librefdrv.so $ROOT/oracle/ref_driver.c $ROOT/oracle/ref_hooks.c $ROOT/oracle/ref_shadow.c \
      -L$D -lhomer_ref -L$ROOT/oracle -loracle -Wl,-rpath,$D -Wl,-rpath,$ROOT/oracle -lpthread -lm -ldl
  HB_REF_DIR=$D python - "$CLIP" <<'PY'
import sys, hashlib, os
root = os.environ.get("HB_ROOT", os.getcwd())
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from _encode import encode, make_yuv
w, h, nf = (int(v) for v in sys.argv[1].split("x"))
bs, rec, _ = encode(w, h, make_yuv(w, h, nf), nf)
print(os.environ["HB_REF_DIR"], len(bs), hashlib.md5(bs).hexdigest())
PY
done
