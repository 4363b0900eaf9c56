#!/bin/bash
for s in 3 4 6; do echo -n "C3 seg $s: "; CRN_TRACE_SEGMENTS=$s python profiles/trace_time.py --config C3 --frames 4 2>&1 | tail -1; done
for s in 1 2 4; do echo -n "C4 seg $s: "; CRN_TRACE_SEGMENTS=$s python profiles/trace_time.py --config C4 --frames 3 2>&1 | tail -1; done
for s in 3 6; do for c in C1 C2; do echo -n "$c seg $s: "; CRN_TRACE_SEGMENTS=$s python profiles/trace_time.py --config $c --frames 6 2>&1 | tail -1; done; done
