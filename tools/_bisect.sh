python integration/make_hor_fasta.py /tmp/c.fa 3 2500 9 2
mkdir -p /tmp/d; CLB_DUMP_DIR=/tmp/d ./oracle/_ref/centrolign_b200 -v 0 -c -y 800 /tmp/c.fa 2>/dev/null | md5sum
ls /tmp/d | wc -l
python tools/check_chain_dump.py /tmp/d 2>&1 | tail -30
