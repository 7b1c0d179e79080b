"""Comparators for bench.py (never imported by legionsolvers_b200): the reference's GPU call sequence through
cuSPARSE / cuBLAS (cusparse_ref).  The CPU reference arm lives in oracle/."""
