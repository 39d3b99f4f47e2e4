for c in "vaihingen_pl 9.0 3" "dales_pl 7.0 2" "vaihingen_pl 16.0 3"; do timeout 600 python tests/ref_dropin_script.py $c 2>&1 | tail -3; done
