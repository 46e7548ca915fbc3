"""nomad_b200: B200-native (sm_100a) implementation of NOMAD's scoring and loss hot path.

``nomad_b200.nomad.Nomad`` mirrors the reference's ``nomad_audio.nomad.Nomad``; the arithmetic lives in
``csrc/`` behind the C ABI of ``include/nomad_b200.h``.  Importing this package does not need a GPU;
constructing ``Nomad`` / ``Engine`` does.
"""
__version__ = "0.1.0"
