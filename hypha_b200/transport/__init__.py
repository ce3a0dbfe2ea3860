"""Control-plane fabric: dial/listen, gossip, DHT records, request/response, tensor streams."""
from .network import (AlreadyRegistered, DialError, HandlerRegistration, IncomingStream, Network,
                      NetworkError, NoHandler, PublishError, RecordNotFound, RequestFailed,
                      StreamAborted, Subscription, new_peer_id)

__all__ = ["AlreadyRegistered", "DialError", "HandlerRegistration", "IncomingStream", "Network",
           "NetworkError", "NoHandler", "PublishError", "RecordNotFound", "RequestFailed",
           "StreamAborted", "Subscription", "new_peer_id"]
