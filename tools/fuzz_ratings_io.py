"""Mutation fuzz of the ratings-file ingest (csrc/ratings_io.cpp, host-only) under AddressSanitizer + UBSan.
usage: python tools/fuzz_ratings_io.py [files=400]
Builds a stand-alone harness (g++ -fsanitize=address,undefined) around mfsgd_read_ratings, mutates a MovieLens-style CSV and a
Netflix-Prize-style file (byte flips, deletions, insertions from an alphabet of digits, separators, quotes, signs, BOM bytes,
NULs; random truncation) and parses every mutant with 1 and 7 slices. Any sanitizer report or crash is printed; exit 1 then.
Round 2: 400 mutants x 2 slice counts, no finding."""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = r'''
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "%s/include/mfsgd.h"
namespace mfsgd { int set_error(int code, const char* fmt, ...) { (void)fmt; return code; } }
int main(int argc, char** argv) {
    for (int a = 1; a < argc; a++) {
        mfsgd_ratings r; memset(&r, 0, sizeof r);
        if (mfsgd_read_ratings(argv[a], 0, &r) == 0) {
            long long s = 0;                                  /* touch everything that came back */
            for (long long t = 0; t < r.n; t++) s += r.users[t] + r.items[t] + (long long)r.ratings[t];
            for (int u = 0; u < r.n_users; u++) s += r.user_ids[u];
            for (int i = 0; i < r.n_items; i++) s += r.item_ids[i];
            if (s == 42) puts("");
            mfsgd_free_ratings(&r);
        }
    }
    return 0;
}
''' % ROOT


def main():
    n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    random.seed(1)
    d = tempfile.mkdtemp(prefix="mfsgd_fuzz_")
    src = os.path.join(d, "harness.cpp")
    open(src, "w").write(HARNESS)
    exe = os.path.join(d, "harness")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-pthread", "-o", exe, src,
                           os.path.join(ROOT, "matrixfactorizationsgd.java_b200", "csrc", "ratings_io.cpp")])
    csv = ("userId,movieId,rating,timestamp\n" + "".join("%d,%d,%g,%d\n" % (random.randint(1, 500), random.randint(1, 300), random.randint(1, 10) / 2,
                                                                         random.randint(10 ** 9, 2 * 10 ** 9)) for _ in range(300))).encode()
    lines = []
    for m in range(1, 30):
        lines.append("%d:" % m)
        lines += ["%d,%d,2005-01-01" % (random.randint(1, 999), random.randint(1, 5)) for _ in range(random.randint(0, 12))]
    netflix = ("\n".join(lines) + "\n").encode()
    alphabet = b"0123456789,.:;-+eE \t\r\n\"'#xyz\xef\xbb\xbf\x00\xff|"
    files = []
    for j in range(n_files):
        b = bytearray(random.choice([csv, netflix]))
        for _ in range(random.randint(1, 12)):
            op, pos = random.random(), random.randrange(len(b))
            if op < 0.4:
                b[pos] = random.choice(alphabet)
            elif op < 0.7:
                del b[pos:pos + random.randint(1, 40)]
            else:
                b[pos:pos] = bytes(random.choice(alphabet) for _ in range(random.randint(1, 8)))
            if not b:
                b = bytearray(b"1")
        if random.random() < 0.3:
            b = b[:random.randrange(1, len(b) + 1)]
        path = os.path.join(d, "f%d.txt" % j)
        open(path, "wb").write(bytes(b))
        files.append(path)
    problems = 0
    for threads in ("1", "7"):
        env = dict(os.environ, MFSGD_IO_THREADS=threads)
        for k in range(0, len(files), 50):
            r = subprocess.run([exe] + files[k:k + 50], capture_output=True, text=True, env=env)
            if r.returncode != 0 or "ERROR" in r.stderr or "runtime error" in r.stderr:
                problems += 1
                print("PROBLEM: slices=%s files %d..%d rc=%d\n%s" % (threads, k, k + 49, r.returncode, r.stderr[-1500:]))
    print("fuzzed %d mutants x 2 slice counts in %s: %d problem batches" % (n_files, d, problems))
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
