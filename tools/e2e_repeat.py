"""Run-to-run spread of the end-to-end pairwise run (two 100 kbp HOR arrays, default options) with the drop-in CLI, for a list of
fill-in pool sizes: python tools/e2e_repeat.py [pool sizes ...] (default: 64 64 64).  Prints wall seconds, md5 of the output and
the library's own call accounting (CLB_COUNT_CALLS)."""
import subprocess, sys, time, os, hashlib
fa = "/tmp/x.fa"
subprocess.run([sys.executable, "integration/make_hor_fasta.py", fa, "2", "100000", "1", "0"], check=True)
for threads in (sys.argv[1:] or ["64", "64", "64"]):
    env = dict(os.environ, CLB_FILL_IN_THREADS=threads, CLB_COUNT_CALLS="1")
    t0 = time.perf_counter()
    r = subprocess.run(["oracle/_ref/centrolign_b200", "-v", "0", fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    dt = time.perf_counter() - t0
    line = [l for l in r.stderr.decode().splitlines() if "chain calls" in l]
    print(f"threads {threads:>4}: {dt:6.2f} s  md5 {hashlib.md5(r.stdout).hexdigest()[:8]}  {line[0][6:] if line else ''}", flush=True)
