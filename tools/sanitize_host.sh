#!/bin/bash
# Host-side memory / UB check (no GPU): the CPU oracle's tests and the native OBJ loader under ASan + UBSan.
# compute-sanitizer is refused on the GPU pool, so this covers what runs on the host; device code is covered by the
# bit-exact comparisons against the oracle.
set -e
cd "$(dirname "$0")/.."
tmp=$(mktemp -d)
echo "== oracle under -fsanitize=address,undefined"
make -C oracle -s
cp oracle/_build/libzdr_oracle.so $tmp/oracle_plain.so
gcc -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -ffp-contract=off -fPIC -shared -I include \
    -o oracle/_build/libzdr_oracle.so oracle/zdr_oracle.c -lm
trap 'cp $tmp/oracle_plain.so oracle/_build/libzdr_oracle.so; touch oracle/_build/libzdr_oracle.so' EXIT
ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    python -m pytest tests/test_oracle_cpu.py tests/test_golden.py -x -q
echo "== zdr_obj_load under -fsanitize=address,undefined"
cat > $tmp/main.cpp <<'CPP'
#include <cstdint>
#include <cstdio>
extern "C" int zdr_obj_load(const char*, float**, int*, int32_t**, int*);
extern "C" void zdr_obj_free(void*);
extern "C" const char* zdr_obj_last_error(void);
int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        float* v; int32_t* t; int nv, nt;
        const int rc = zdr_obj_load(argv[i], &v, &nv, &t, &nt);
        printf("%-40s rc=%d nv=%d nt=%d %s\n", argv[i], rc, nv, nt, rc ? zdr_obj_last_error() : "");
        if (!rc) { zdr_obj_free(v); zdr_obj_free(t); }
    }
}
CPP
g++ -std=c++17 -g -O1 -fsanitize=address,undefined -fno-omit-frame-pointer -x c++ zdr_b200/csrc/obj_loader.cu $tmp/main.cpp -o $tmp/obj_asan
mkdir -p $tmp/objs
printf 'v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3' > $tmp/objs/no_newline.obj
printf 'v 0 0 0\r\nv 1 0 0\r\nv 0 1 0\r\nf -1 -2 0\r\n' > $tmp/objs/crlf_negative.obj
printf 'v 0 0 0\nf 1 2 4\n' > $tmp/objs/out_of_range.obj
printf 'v 0 0 zero\nf 1 1 1\n' > $tmp/objs/bad_number.obj
printf 'f\nv\n\n#\nv 1 2 3\nf 1\nf 1 1 1 1 1\n' > $tmp/objs/fragments.obj
printf 'v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvn 0 0 1\nf 1/1/1 2//1 3/1\nf 1// 2/ 3\n' > $tmp/objs/slashes.obj
printf 'v 0 0 0\nf 99999999999999999999 1 1\n' > $tmp/objs/huge_index.obj
head -c 5000 /dev/urandom > $tmp/objs/binary.obj
: > $tmp/objs/empty.obj
$tmp/obj_asan $tmp/objs/*.obj zdr_b200/assets/*.obj /nonexistent.obj
echo "clean"
