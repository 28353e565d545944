"""federated_multi_modal_b200 — B200-native MaPLe forward/backward step + FedAvg hot path.

Layout: ``csrc/`` CUDA kernels + C ABI (libmfk.so, declared in include/mfk.h), ``_lib``/``ops`` ctypes
binding, ``engine`` explicit fwd/bwd schedule, ``clip/`` and ``trainers/`` the host-side mirror of the
reference's plugin interface (same class names, signatures and state_dict keys), ``fed`` the round-end
exchange, ``synth`` seeded synthetic weights/inputs. Importing the package does not need a GPU; running
any op does (there is no CPU fallback).
"""
__version__ = "0.1.0"
