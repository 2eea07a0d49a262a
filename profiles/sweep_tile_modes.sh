for nq in 1e8 5e7 3e7 1.5e7; do for m in a b c d; do echo "nq=$nq mode=$m $(IPX_TILE_MODE=$m python profiles/prof_eval3d.py $nq f64 256 | grep 'sorted rep2')"; done; done
