class Slider:
    pass


class Button:
    pass
