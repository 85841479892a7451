"""ORACLE package — CPU restatements of the reference used ONLY as the checker by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
