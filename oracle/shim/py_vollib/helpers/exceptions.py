class PriceIsAboveMaximum(Exception):
    def __init__(self):
        Exception.__init__(self, "The volatility is above the maximum price.")


class PriceIsBelowIntrinsic(Exception):
    def __init__(self):
        Exception.__init__(self, "The volatility is below the intrinsic value.")
