from xlb_b200.velocity_set.velocity_set import VelocitySet
from xlb_b200.velocity_set.d2q9 import D2Q9
from xlb_b200.velocity_set.d3q19 import D3Q19
from xlb_b200.velocity_set.d3q27 import D3Q27
