class SimSnap:
    def __init__(self, **arrays):
        self._a = dict(arrays)

    def __getitem__(self, k):
        return self._a[k]

    def __len__(self):
        return len(self._a["mass"])
