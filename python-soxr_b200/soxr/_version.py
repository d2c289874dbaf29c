version = "0.1.0+b200"
