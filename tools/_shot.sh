cd /root/repo
timeout 600 python tools/_exp_f32tc.py 2>&1 | grep -v Warn | tail -30
