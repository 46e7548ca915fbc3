"""``nomad_audio.nomad`` submodule of the reference (``src/nomad_audio/nomad.py``): ``from nomad_audio.nomad import
Nomad`` works as it does there.  The implementation lives in ``nomad_b200.nomad``."""
from nomad_b200.nomad import (Nomad, NomadLoss, LossNetLayers, TripletModel, nomad_path, w2v_path)  # noqa: F401
