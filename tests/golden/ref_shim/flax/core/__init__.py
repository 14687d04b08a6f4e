def freeze(x):
  return x


def unfreeze(x):
  return x
