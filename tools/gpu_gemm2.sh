timeout 300 python tools/probe_gemm2.py 2>&1 | grep '"mode": "2"'
timeout 200 python tools/probe_multinomial.py cfg5 2>&1 | tail -1 | cut -c1-260
