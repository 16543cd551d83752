cd /root/repo
bash tools/r02_full.sh
