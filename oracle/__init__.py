"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's dense-inference hot path (GIGA
`LocalVoxelEncoder` -> shared tri-plane U-Net -> `LocalDecoder` heads).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import anything from this package, and there only as
the checker / the timed CPU baseline -- never from `giga_b200/` (the product).
"""
