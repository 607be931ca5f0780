"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy) restating the Remhos RK-stage hot path.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  See DESIGN.md for what it is pinned against."""
