python integration/make_hor_fasta.py /tmp/x.fa 2 100000 1 0
for i in 1 2; do
CLB_FILL_IN_THREADS=1 CLB_TIMING=1 CLB_COUNT_CALLS=1 oracle/_ref/centrolign_b200 -v 3 /tmp/x.fa 2>/tmp/err.txt | md5sum
grep "\] chain: " /tmp/err.txt | awk '{ if ($4+0 > 50 || $14+0 > 50) print }' | cut -c1-330
grep "elapsed\|affine chaining\|chain calls\|create\b" /tmp/err.txt | grep -v "create:" | cut -c1-200
done
