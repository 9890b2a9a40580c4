from .embedded_graph import *
from .josephson_circuit import *
from .current_phase_relation import *
from .time_evolution import *
from .sources import RankOneSource
