from xlb_b200.distribute.distribute import distribute, distribute_operator
