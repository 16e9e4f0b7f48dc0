class HumanOutputFormat(object):
    def __init__(self, *a, **k):
        pass

    def writekvs(self, kvs):
        pass
