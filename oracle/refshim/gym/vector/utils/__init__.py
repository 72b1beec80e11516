from ... import spaces
