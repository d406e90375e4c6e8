"""placeholder — replaced below"""
