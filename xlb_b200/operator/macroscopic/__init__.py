from xlb_b200.operator.macroscopic.macroscopic import Macroscopic
from xlb_b200.operator.macroscopic.second_moment import SecondMoment
from xlb_b200.operator.macroscopic.zero_moment import ZeroMoment
from xlb_b200.operator.macroscopic.first_moment import FirstMoment
