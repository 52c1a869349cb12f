"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the D3Feat hot path.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs only.  Nothing under d3feat/ may import this package.
"""
