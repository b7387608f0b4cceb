"""``Network`` container (reference: ``code/network_class.py:2-21``); pickled as
``{dataset}_{network}.network.pickle`` (``importance_scores.py:386-415``), so attribute names
are part of the file format."""


class Network:
    def __init__(self, name):
        self.name = name
        self.rulesets = None
        self.nodes = None
        self.cells = None
        self.network = None
        self.dataset = None

        self.attractors = None
        self.importance_score = None

        self.representative_attractors = {}
        self.filtered_attractors = None
        self.filtered_attractor_indices = None
        self.cell_map = {}
