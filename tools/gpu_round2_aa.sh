cd $GRAFT_REPO_ROOT
timeout 300 python tools/prof_host_step.py 4 2>&1 | grep -v "^$" | cut -c1-170 | head -110
