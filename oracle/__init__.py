"""ORACLE package — test infrastructure, NOT product code (see ref_port.py / xm3d_oracle.c)."""
