"""CPU oracle for the hot path — TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is imported by the product package (`two-tower-model-v2_b200/`).  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import or execute it, and only as the checker or the reported CPU baseline.
"""
