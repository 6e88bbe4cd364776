"""ORACLE: CPU restatement of the reference's hot path.  Test infrastructure only --
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never from the product package `typlonk_b200`."""
