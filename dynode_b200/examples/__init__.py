"""The reference's example models (examples/*.py) re-created on the B200 engine.

Each module keeps the reference's names (get_config, get_odeparams, <model>_ode, *_ODEParams) so
user code and tests read the same; the right-hand sides are ordinary torch functions registered
with `@flow_family`, which is what lets `simulate` run them on the device kernels.
"""
