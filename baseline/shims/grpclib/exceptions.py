class GRPCError(Exception):
    def __init__(self, status, message=None, details=None):
        super().__init__(status, message, details)
        self.status = status
        self.message = message


class ProtocolError(Exception):
    pass


class StreamTerminatedError(Exception):
    pass
