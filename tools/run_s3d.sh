for d in 0 1 2 4 3 7; do echo "dbg=$d"; SHGAN_HALO_DBG=$d timeout 200 python tools/halo_profile.py 2>&1 | tail -3 | cut -c1-400; done
