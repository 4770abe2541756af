// stand-in: nothing of this header is used by the translation units compiled under oracle/_ref
