class ResultsWriter(object):
    def __init__(self, *a, **k):
        pass

    def write_row(self, epinfo):
        pass
