// stand-in: not used by the translation units compiled under oracle/_ref
