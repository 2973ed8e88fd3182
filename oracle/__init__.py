"""ORACLE — test infrastructure only (parity unpinned; see oracle/orc_field.hpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
