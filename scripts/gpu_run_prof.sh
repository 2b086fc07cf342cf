for env in "HN_NO_PAIR=1" "HN_PAIR=1"; do env $env timeout 180 python scripts/profile_halo_pair.py 2>&1 | tail -6; done
