from xlb_b200.operator.operator import Operator
from xlb_b200.operator.parallel_operator import ParallelOperator
