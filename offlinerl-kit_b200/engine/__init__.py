"""Host-side engine: arena, descriptor builders, launch plans and per-algorithm step schedules.

The schedules are Python (the reference is Python); every launch goes through the C ABI of
liborlk_b200.so and is replayed as one CUDA graph per gradient step.
"""
