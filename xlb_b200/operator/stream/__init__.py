from xlb_b200.operator.stream.stream import Stream
