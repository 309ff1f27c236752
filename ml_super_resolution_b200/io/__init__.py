"""Wire formats of the reference's data files: TFRecord framing + tf.train.Example patch pairs."""
