from xlb_b200.operator.collision.collision import Collision
from xlb_b200.operator.collision.bgk import BGK
from xlb_b200.operator.collision.kbc import KBC
