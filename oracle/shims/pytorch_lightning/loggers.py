"""ORACLE SHIM: placeholder sub-module (only scripts/run.py uses it; the CLI is
out of scope)."""
